"""Hyperparameter-optimization config (reference: maggy/config/hyperparameter_optimization.py:33-93)."""
from __future__ import annotations

from typing import Optional, Union

from ..earlystop import AbstractEarlyStop
from ..optimizer import AbstractOptimizer
from ..searchspace import Searchspace
from .lagom import LagomConfig


class HyperparameterOptConfig(LagomConfig):
    """
    :param num_trials: number of trials to evaluate (pruners / grid search derive their own count)
    :param optimizer: "randomsearch" | "asha" | "tpe" | "gp" | "gridsearch" | "none" | AbstractOptimizer
    :param searchspace: a :class:`Searchspace`
    :param optimization_key: key of the metric when train_fn returns a dict
    :param direction: "max" or "min"
    :param es_interval: check early stopping every ``es_interval`` reported steps
    :param es_min: only start stopping once more than ``es_min`` trials have finished
    :param es_policy: "median" | "none" | AbstractEarlyStop
    :param model / dataset: optional objects injected into train_fn by name
    :param gpus_per_trial: NEW (not in the reference): size of the GPU gang each trial runs on
        (BASELINE config 4: "4 trials x 2 GPUs")
    """

    def __init__(self, num_trials: int, optimizer: Union[str, AbstractOptimizer], searchspace: Searchspace,
                 optimization_key: str = "Metric", direction: str = "max", es_interval: int = 1,
                 es_min: int = 10, es_policy: Union[str, AbstractEarlyStop] = "median",
                 name: str = "HPOptimization", description: str = "", hb_interval: int = 1,
                 model=None, dataset=None, gpus_per_trial: int = 1, num_workers: Optional[int] = None,
                 seed: Optional[int] = None, resume_from: Optional[str] = None):
        """``resume_from`` (additive, not in the reference — SURVEY.md §5.4): log directory of an earlier run of the
        same experiment; trials whose ``<trial_id>/trial.json`` is FINALIZED there are not executed again — when the
        controller suggests the same configuration its saved result is replayed into the stores."""
        super().__init__(name, description, hb_interval)
        if not isinstance(num_trials, int) or isinstance(num_trials, bool) or num_trials <= 0:
            raise ValueError(f"Number of trials should be greater than zero! got {num_trials!r}")
        if str(direction).lower() not in ("max", "min"):
            raise ValueError(f"The experiment's direction should be a string (either 'min' or 'max') but it is {direction!r}")
        self.num_trials = num_trials
        self.optimizer = optimizer
        self.optimization_key = optimization_key
        self.searchspace = searchspace
        self.direction = str(direction).lower()
        self.es_policy = es_policy
        self.es_interval = es_interval
        self.es_min = es_min
        self.model = model
        self.dataset = dataset
        self.gpus_per_trial = gpus_per_trial
        self.num_workers = num_workers
        self.seed = seed
        self.resume_from = resume_from


# the docs / notebooks of the reference also use this name (docs/hpo/strategies.md:20)
OptimizationConfig = HyperparameterOptConfig
