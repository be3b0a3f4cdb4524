"""Asynchronous Successive Halving (reference: maggy/optimizer/asha.py:23-169; Li et al. 2018).

Rung ``k`` runs configurations with ``budget = resource_min * reduction_factor**k``. Whenever a worker
asks for work, the highest rung with a promotable configuration (one of the top ``1/reduction_factor``
finished trials of that rung that has not been promoted yet) wins; otherwise a fresh random
configuration starts in rung 0. The experiment ends once a configuration has been placed in the top
rung. Unlike the reference (which always ranks descending, asha.py:161-169) the ranking honours the
experiment's ``direction``.
"""
import math

from ..trial import Trial
from .abstractoptimizer import AbstractOptimizer


class Asha(AbstractOptimizer):
    def __init__(self, reduction_factor=2, resource_min=1, resource_max=4):
        super().__init__()
        if not isinstance(reduction_factor, int) or isinstance(reduction_factor, bool) or reduction_factor < 2:
            raise Exception("Can't initialize ASHA optimizer. 'reduction_factor' has to be an integer equal to "
                            f"or larger than 2: {reduction_factor}")
        if not isinstance(resource_min, int) or isinstance(resource_min, bool):
            raise Exception("Can't initialize ASHA optimizer. 'resource_min' not of type INTEGER.")
        if not isinstance(resource_max, int) or isinstance(resource_max, bool):
            raise Exception("Can't initialize ASHA optimizer. 'resource_max' not of type INTEGER.")
        if resource_min >= resource_max:
            raise Exception("Can't initialize ASHA optimizer. 'resource_min' is larger than 'resource_max'.")
        self.reduction_factor = reduction_factor
        self.resource_min = resource_min
        self.resource_max = resource_max

    def initialize(self):
        self.rungs = {0: []}      # rung -> trials placed there
        self.promoted = {0: []}   # rung -> ids already promoted out of it
        self.max_rung = int(math.floor(math.log(self.resource_max / self.resource_min, self.reduction_factor) + 1e-9))
        assert self.num_trials >= self.reduction_factor ** (self.max_rung + 1), (
            f"ASHA needs num_trials >= reduction_factor**(max_rung+1) = "
            f"{self.reduction_factor ** (self.max_rung + 1)}, got {self.num_trials}")

    def budget_of(self, rung: int) -> int:
        return self.resource_min * self.reduction_factor ** rung

    def _finished(self, rung):
        return [t for t in self.rungs.get(rung, []) if t.status == Trial.FINALIZED]

    def _top_k(self, rung, k):
        if k <= 0:
            return []
        done = self._finished(rung)
        done.sort(key=lambda t: t.final_metric, reverse=(self.direction != "min"))
        return done[:k]

    def _promotion(self):
        for k in range(self.max_rung - 1, -1, -1):
            quota = len(self._finished(k)) // self.reduction_factor
            already = self.promoted.setdefault(k, [])
            if quota - len(already) <= 0:
                continue
            fresh = [t for t in self._top_k(k, quota) if t.trial_id not in already]
            if not fresh:
                continue
            parent = fresh[0]
            params = dict(parent.params)
            params["budget"] = self.budget_of(k + 1)
            child = Trial(params, info_dict={"run_budget": params["budget"], "sample_type": "promoted",
                                             "parent": parent.trial_id})
            self.rungs.setdefault(k + 1, []).append(child)
            already.append(parent.trial_id)
            self._log(f"promote {parent.trial_id} from rung {k} to rung {k + 1} as {child.trial_id}")
            return child
        return None

    # ---- device-side decisions (K9, csrc/include/mg_asha.h) ------------------------------------------------------
    NONE, NEW, PROMOTE, NEW_PUBLISHED = 0, 1, 2, 3

    def rung_of(self, trial) -> int:
        budget = trial.params.get("budget", self.resource_min)
        return int(round(math.log(max(budget, 1) / self.resource_min, self.reduction_factor)))

    def _child_of(self, parent, rung):
        params = dict(parent.params)
        params["budget"] = self.budget_of(rung + 1)
        child = Trial(params, info_dict={"run_budget": params["budget"], "sample_type": "promoted",
                                         "parent": parent.trial_id})
        self.rungs.setdefault(rung + 1, []).append(child)
        self.promoted.setdefault(rung, []).append(parent.trial_id)
        return child

    def apply_device_decision(self, kind, rung, parent_id, trial=None):
        """Materialise a next-trial decision the persistent scheduler kernel took from its own rung tables: the driver
        only copies / samples the hyperparameters. Falls back to :meth:`get_suggestion` when the decision cannot be
        honoured (unknown parent, or the host already promoted it after a missed decision)."""
        if self.max_rung in self.rungs:
            return None
        if kind == self.PROMOTE:
            parent = next((t for t in self.rungs.get(rung, []) if t.trial_id == parent_id), None)
            if parent is None or parent.trial_id in self.promoted.setdefault(rung, []):
                return self.get_suggestion(trial)
            child = self._child_of(parent, rung)
            self._log(f"[device] promote {parent.trial_id} from rung {rung} to rung {rung + 1} as {child.trial_id}")
            return child
        if kind != self.NEW:
            return self.get_suggestion(trial)
        return self.fresh_trial()

    def fresh_trial(self):
        """A new random configuration in rung 0 (also what the driver parks as a worker's prefetched next trial)."""
        params = self.searchspace.get_random_parameter_values(1)[0]
        params["budget"] = self.resource_min
        fresh = Trial(params, info_dict={"run_budget": self.resource_min, "sample_type": "random"})
        self.rungs[0].append(fresh)
        return fresh

    def get_suggestion(self, trial=None):
        if trial is not None:
            if self.max_rung in self.rungs:
                return None  # a configuration reached the top rung: wind the experiment down
            child = self._promotion()
            if child is not None:
                return child
        params = self.searchspace.get_random_parameter_values(1)[0]
        params["budget"] = self.resource_min
        fresh = Trial(params, info_dict={"run_budget": self.resource_min, "sample_type": "random"})
        self.rungs[0].append(fresh)
        return fresh

    def finalize_experiment(self, trials):
        return
