"""Never stop (reference: maggy/earlystop/nostop.py:20-25)."""
from .abstractearlystop import AbstractEarlyStop


class NoStoppingRule(AbstractEarlyStop):
    @staticmethod
    def earlystop_check(to_check, finalized_trials, direction):
        return None
