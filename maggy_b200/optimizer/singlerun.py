"""Controller for experiments without a search space: ``num_trials`` runs with empty hyperparameters
(reference: maggy/optimizer/singlerun.py:21-37). Repeated runs get a ``run`` index so their trial ids
(and artefact directories) do not collide."""
from ..trial import Trial
from .abstractoptimizer import AbstractOptimizer


class SingleRun(AbstractOptimizer):
    def __init__(self):
        super().__init__()
        self.trial_buffer = []

    def initialize(self):
        n = self.num_trials
        self.trial_buffer = [Trial({} if n == 1 else {"run": i}) for i in reversed(range(n))]

    def get_suggestion(self, trial=None):
        return self.trial_buffer.pop() if self.trial_buffer else None

    def finalize_experiment(self, trials):
        return
