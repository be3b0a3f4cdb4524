"""Base class of all search controllers (reference: maggy/optimizer/abstractoptimizer.py:28-442).

The driver injects ``searchspace``, ``num_trials``, ``trial_store`` (running trials, id -> Trial),
``final_store`` (finished trials, list) and ``direction`` before calling ``_initialize``. A controller
answers ``get_suggestion(trial)`` with a new :class:`Trial`, the string ``"IDLE"`` (nothing to schedule
right now, ask again) or ``None`` (experiment exhausted).
"""
from __future__ import annotations

import os
import time
from abc import ABC, abstractmethod
from datetime import datetime

import numpy as np

from ..core.environment.singleton import EnvSing
from ..pruner.hyperband import Hyperband
from ..trial import Trial

_SAMPLE_TYPES = ("random", "random_forced", "model", "promoted", "grid")
_PRUNERS = {"hyperband": Hyperband}


class AbstractOptimizer(ABC):
    def __init__(self, pruner=None, pruner_kwargs=None):
        self.searchspace = None
        self.num_trials = None
        self.trial_store = None
        self.final_store = None
        self.direction = None
        self.pruner = None
        if pruner:
            self.init_pruner(pruner, pruner_kwargs)
        self.log_file = None
        self.fd = None
        self.sampling_time_start = 0.0

    # ------------------------------------------------------------------ to implement
    @abstractmethod
    def initialize(self):
        """Called once after the driver injected the shared state."""

    @abstractmethod
    def get_suggestion(self, trial=None):
        """``trial`` is the trial that just finished on the asking worker (or None at start)."""

    @abstractmethod
    def finalize_experiment(self, trials):
        """Called once with all finished trials."""

    def name(self):
        return type(self).__name__

    # ------------------------------------------------------------------ lifecycle used by the driver
    def _initialize(self, exp_dir):
        self._initialize_logger(exp_dir)
        self.initialize()
        self._log(f"Initilized Optimizer {self.name()}")
        if self.pruner:
            self.pruner.initialize_logger(exp_dir=exp_dir)

    def _finalize_experiment(self, trials):
        self.finalize_experiment(trials)
        self._log("Experiment finished")
        self._close_log()
        if self.pruner:
            self.pruner._close_log()

    def _initialize_logger(self, exp_dir):
        env = EnvSing.get_instance()
        self.log_file = os.path.join(exp_dir, "optimizer.log")
        if not env.exists(self.log_file):
            env.dump("", self.log_file)
        self.fd = env.open_file(self.log_file, flags="w")
        self._log("Initialized Optimizer Logger")

    def _log(self, msg):
        if self.fd and not self.fd.closed:
            self.fd.write(EnvSing.get_instance().str_or_byte(f"{datetime.now().isoformat()}: {msg}\n"))

    def _close_log(self):
        if self.fd and not self.fd.closed:
            self.fd.flush()
            self.fd.close()

    # ------------------------------------------------------------------ views on finished trials
    @staticmethod
    def _selector(trial_ids):
        if isinstance(trial_ids, str) and trial_ids == "all":
            return lambda tid: True
        if isinstance(trial_ids, str):
            return lambda tid: tid == trial_ids
        ids = set(trial_ids)
        return lambda tid: tid in ids

    def _budget_match(self, trial, budget):
        return budget in (0, None) or trial.params.get("budget") == budget

    def get_hparams_dict(self, trial_ids="all"):
        keep = self._selector(trial_ids)
        return {t.trial_id: t.params for t in self.final_store if keep(t.trial_id)}

    def get_hparams_array(self, budget=0):
        return np.array([self.searchspace.dict_to_list(t.params) for t in self.final_store
                         if self._budget_match(t, budget)])

    def get_metrics_dict(self, trial_ids="all"):
        """``{trial_id: metric}`` of finished trials, sign-flipped so that LOWER IS BETTER."""
        sign = -1 if self.direction == "max" else 1
        keep = self._selector(trial_ids)
        return {t.trial_id: t.final_metric * sign for t in self.final_store if keep(t.trial_id)}

    def get_metrics_array(self, budget=0, interim_metrics=False):
        vals = [np.array(t.metric_history) if interim_metrics else t.final_metric
                for t in self.final_store if self._budget_match(t, budget)]
        arr = np.array(vals, dtype=object if interim_metrics else float)
        return -arr if self.direction == "max" else arr

    def hparams_exist(self, trial) -> bool:
        """True if the same configuration (ignoring ``budget``) already ran or is running."""
        names = self.searchspace.keys()

        def core(params):
            return {k: params[k] for k in names if k in params}

        mine = core(trial.params)
        for idx, done in enumerate(self.final_store):
            if core(done.params) == mine:
                self._log(f"WARNING Duplicate Config: Hparams {trial.params} are equal to params of finished "
                          f"trial no. {idx}: {done.trial_id}")
                return True
        for busy in list(self.trial_store.values()):
            if core(busy.params) == mine:
                self._log(f"WARNING Duplicate Config: Hparams {trial.params} are equal to currently evaluating "
                          f"Trial: {busy.trial_id}")
                return True
        return False

    # ------------------------------------------------------------------ pruner / trial creation
    def init_pruner(self, pruner, pruner_kwargs):
        if pruner not in _PRUNERS:
            raise ValueError(f"expected pruner to be in {list(_PRUNERS)}, got {pruner}")
        self.pruner = _PRUNERS[pruner](trial_metric_getter=self.get_metrics_dict, **(pruner_kwargs or {}))

    def create_trial(self, hparams, sample_type, run_budget=0, model_budget=None):
        if sample_type not in _SAMPLE_TYPES:
            raise ValueError(f"expected sample_type to be in {list(_SAMPLE_TYPES)}, got {sample_type}")
        if sample_type == "model" and model_budget is None:
            raise ValueError("expected `model_budget` because sample_type==`model`, got None")
        info = {"run_budget": run_budget, "sample_type": sample_type,
                "sampling_time": time.time() - self.sampling_time_start}
        self.sampling_time_start = 0.0
        if model_budget is not None:
            info["model_budget"] = model_budget
        if run_budget > 0:
            hparams["budget"] = run_budget
        return Trial(hparams, trial_type="optimization", info_dict=info)

    def get_max_budget(self):
        if self.pruner:
            return self.pruner.max_budget
        if len(self.final_store) == 0:
            raise ValueError("At least one finalized Trial is necessary to calculate max budget")
        return len(self.final_store[0].metric_history)

    def ybest(self, budget=0):
        return np.min(self.get_metrics_array(budget=budget))

    def yworst(self, budget=0):
        return np.max(self.get_metrics_array(budget=budget))

    def ymean(self, budget=0):
        return np.mean(self.get_metrics_array(budget=budget))
