"""Random search, optionally under a Hyperband pruner (reference: maggy/optimizer/randomsearch.py:23-113)."""
import time
from copy import deepcopy

from ..searchspace import Searchspace
from .abstractoptimizer import AbstractOptimizer


class RandomSearch(AbstractOptimizer):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.config_buffer = []

    def initialize(self):
        kinds = set(self.searchspace.names().values())
        if not kinds & {Searchspace.DOUBLE, Searchspace.INTEGER}:
            raise NotImplementedError("Searchspace needs at least one continuous parameter for random search.")
        # all configurations are drawn up front (so the sequence only depends on the RNG state at start)
        self.config_buffer = self.searchspace.get_random_parameter_values(self.num_trials)

    def _pruned_suggestion(self):
        run = self.pruner.pruning_routine()
        if run == "IDLE":
            self._log("Worker is IDLE and has to wait until a new trial can be scheduled")
            return "IDLE"
        if run is None:
            self._log("Experiment has finished")
            return None
        parent = run["trial_id"]
        if parent:
            hparams = deepcopy(self.get_hparams_dict(trial_ids=parent)[parent])
            trial = self.create_trial(hparams=hparams, sample_type="promoted", run_budget=run["budget"])
            self._log(f"use hparams from promoted trial {parent}")
        else:
            hparams = self.searchspace.get_random_parameter_values(1)[0]
            trial = self.create_trial(hparams=hparams, sample_type="random", run_budget=run["budget"])
        self.pruner.report_trial(original_trial_id=parent, new_trial_id=trial.trial_id)
        return trial

    def get_suggestion(self, trial=None):
        self._log("### start get_suggestion ###")
        self.sampling_time_start = time.time()
        if self.pruner:
            nxt = self._pruned_suggestion()
        elif self.config_buffer:
            nxt = self.create_trial(hparams=self.config_buffer.pop(), sample_type="random", run_budget=0)
        else:
            return None
        if nxt not in (None, "IDLE"):
            self._log(f"start trial {nxt.trial_id}: {nxt.params}, {nxt.info_dict} \n")
        return nxt

    def finalize_experiment(self, trials):
        return
