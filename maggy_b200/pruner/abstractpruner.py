"""Pruner interface (reference: maggy/pruner/abstractpruner.py:23-95)."""
from __future__ import annotations

import os
from abc import ABC, abstractmethod
from datetime import datetime

from ..core.environment.singleton import EnvSing


class AbstractPruner(ABC):
    def __init__(self, trial_metric_getter):
        """``trial_metric_getter(ids)`` -> ``{trial_id: metric}`` for finished trials (lower = better):
        the optimizer's ``get_metrics_dict``."""
        self.trial_metric_getter = trial_metric_getter
        self.log_file = None
        self.fd = None

    @abstractmethod
    def pruning_routine(self):
        """-> ``{"trial_id": None | id, "budget": b}`` | ``"IDLE"`` | ``None`` (experiment finished)."""

    @abstractmethod
    def report_trial(self, original_trial_id, new_trial_id):
        """Tell the pruner which trial object was started for the run it handed out last."""

    @abstractmethod
    def finished(self) -> bool:
        ...

    @abstractmethod
    def num_trials(self) -> int:
        ...

    def name(self):
        return type(self).__name__

    def initialize_logger(self, exp_dir):
        env = EnvSing.get_instance()
        self.log_file = os.path.join(exp_dir, "pruner.log")
        if not env.exists(self.log_file):
            env.dump("", self.log_file)
        self.fd = env.open_file(self.log_file, flags="w")
        self._log("Initialized Pruner Logger")

    def _log(self, msg):
        if self.fd and not self.fd.closed:
            self.fd.write(EnvSing.get_instance().str_or_byte(f"{datetime.now().isoformat()}: {msg}\n"))

    def _close_log(self):
        if self.fd and not self.fd.closed:
            self.fd.flush()
            self.fd.close()
