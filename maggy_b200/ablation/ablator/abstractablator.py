"""Ablator interface (reference: maggy/ablation/ablator/abstractablator.py:20-86)."""
from abc import ABC, abstractmethod


class AbstractAblator(ABC):
    def __init__(self, ablation_study, final_store):
        self.ablation_study = ablation_study
        self.final_store = final_store
        self.trial_buffer = []

    @abstractmethod
    def get_number_of_trials(self):
        """How many trials the study consists of (the driver sizes the experiment with it)."""

    @abstractmethod
    def get_dataset_generator(self, ablated_feature, dataset_type="tfrecord"):
        """Callable producing the dataset without ``ablated_feature``."""

    @abstractmethod
    def get_model_generator(self, layer_identifier):
        """Callable producing the model without the identified layer(s)."""

    @abstractmethod
    def initialize(self):
        """Prepare the trial buffer."""

    @abstractmethod
    def get_trial(self, ablation_trial=None):
        """Next Trial, or None when the study is exhausted."""

    @abstractmethod
    def finalize_experiment(self, trials):
        ...

    def name(self):
        return type(self).__name__
