"""Exhaustive grid over DISCRETE / CATEGORICAL hyperparameters (reference: maggy/optimizer/gridsearch.py:23-92)."""
import itertools
import time

from ..searchspace import Searchspace
from .abstractoptimizer import AbstractOptimizer


class GridSearch(AbstractOptimizer):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.config_buffer = []

    def initialize(self):
        self._validate_searchspace(self.searchspace)
        self.config_buffer = self._grid_params(self.searchspace)

    @classmethod
    def get_num_trials(cls, searchspace) -> int:
        cls._validate_searchspace(searchspace)
        n = 1
        for item in searchspace.items():
            n *= len(item["values"])
        return n

    def get_suggestion(self, trial=None):
        if self.pruner:
            raise NotImplementedError("Grid search in combination with trial pruning is currently not supported.")
        if not self.config_buffer:
            return None
        self.sampling_time_start = time.time()
        nxt = self.create_trial(hparams=self.config_buffer.pop(), sample_type="grid", run_budget=0)
        self._log(f"start trial {nxt.trial_id}: {nxt.params}, {nxt.info_dict} \n")
        return nxt

    def finalize_experiment(self, trials):
        return

    @staticmethod
    def _grid_params(searchspace):
        axes = [item["values"] for item in searchspace.items()]
        return [searchspace.list_to_dict(list(combo)) for combo in itertools.product(*axes)]

    @staticmethod
    def _validate_searchspace(searchspace):
        kinds = set(searchspace.names().values())
        if kinds & {Searchspace.DOUBLE, Searchspace.INTEGER}:
            raise NotImplementedError(
                "Searchspace can only contain `discrete` or `categorical` hyperparameters for grid search.")
