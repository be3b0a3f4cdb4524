"""Early-stopping policy interface (reference: maggy/earlystop/abstractearlystop.py:20-42)."""
from abc import ABC, abstractmethod


class AbstractEarlyStop(ABC):
    @staticmethod
    @abstractmethod
    def earlystop_check(to_check, finalized_trials, direction):
        """Return ``to_check.trial_id`` if the running trial should stop, else ``None``."""
