"""Hyperband: parallel successive-halving brackets (reference: maggy/pruner/hyperband.py:29-594).

Same schedule as the reference (golden tables in tests: ``min=1,max=9,eta=3,n_iterations=3`` ->
budgets ``[1,3,9]``, brackets ``[9,3,1] / [3,1] / [3]``, 20 trials). The implementation is organised
around two small classes: a ``_Rung`` (capacity, budget, the runs placed in it) and a ``Bracket`` (one
successive-halving run = a list of rungs and a cursor); ``Hyperband`` starts brackets lazily so that
workers that would otherwise idle at a rung barrier begin the next bracket.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

from .abstractpruner import AbstractPruner


@dataclass
class _Slot:
    origin: str                 # id of the rung-0 trial this configuration descends from
    started: Optional[str] = None   # id of the trial actually started in this rung


@dataclass
class _Rung:
    capacity: int
    budget: int
    handed_out: int = 0
    slots: List[_Slot] = field(default_factory=list)


class Bracket:
    """One successive-halving run."""

    INIT, RUNNING, FINISHED = "INIT", "RUNNING", "FINISHED"

    def __init__(self, n_configs: List[int], budgets: List[int], iteration_id: int,
                 trial_metric_getter: Callable, logger: Callable):
        self.iteration_id = iteration_id
        self.state = Bracket.INIT
        self.n_configs = list(n_configs)
        self.budgets = list(budgets)
        self.rungs = [_Rung(c, b) for c, b in zip(n_configs, budgets)]
        self.current_rung = 0
        self._metrics = trial_metric_getter
        self._log = logger

    # reference attribute names, used by tests / logs
    @property
    def n_rungs(self):
        return len(self.rungs)

    @property
    def actual_n_configs(self):
        return [r.handed_out for r in self.rungs]

    def _done(self, trial_id) -> bool:
        return bool(trial_id) and bool(self._metrics(trial_id))

    def get_next_run(self):
        rung = self.rungs[self.current_rung]
        if rung.handed_out > rung.capacity:
            raise ValueError(f"Too many configs have been sampled in iteration {self.iteration_id}")
        if rung.handed_out < rung.capacity:
            if self.current_rung == 0:
                rung.handed_out += 1
                return {"trial_id": None, "budget": rung.budget}
            for slot in rung.slots:
                if slot.started is None:
                    rung.handed_out += 1
                    return {"trial_id": slot.origin, "budget": rung.budget}
            return None
        if self.promotable():
            self.promote()
            return self.get_next_run()
        if self.finished():
            self.state = Bracket.FINISHED
            self._log(f"{self.iteration_id}. Iteration finished")
        return None

    def report_trial(self, original_trial_id, new_trial_id):
        rung = self.rungs[self.current_rung]
        if self.current_rung == 0:
            rung.slots.append(_Slot(origin=new_trial_id, started=new_trial_id))
        else:
            for slot in rung.slots:
                if slot.origin == original_trial_id and slot.started is None:
                    slot.started = new_trial_id
                    break
        self._log(f"{self.iteration_id}. Iteration, {self.current_rung}. Rung. "
                  f"Started Trial {rung.handed_out}/{rung.capacity}")

    def promotable(self) -> bool:
        rung = self.rungs[self.current_rung]
        if len(rung.slots) < rung.capacity or self.current_rung == self.n_rungs - 1:
            return False
        return all(self._done(s.started) for s in rung.slots)

    def promote(self):
        rung = self.rungs[self.current_rung]
        scores: Dict[str, float] = self._metrics([s.started for s in rung.slots])
        ranked = sorted(scores, key=scores.get)  # getter yields "lower is better"
        nxt = self.rungs[self.current_rung + 1]
        winners = ranked[:nxt.capacity]
        nxt.slots = [_Slot(origin=t) for t in winners]
        self._log(f"{self.iteration_id}. Iteration finished rung: {self.current_rung} \n with trials: {ranked} \n"
                  f" promoted trials: {winners}")
        self.current_rung += 1

    def finished(self) -> bool:
        rung = self.rungs[self.current_rung]
        return (self.current_rung == self.n_rungs - 1 and len(rung.slots) >= rung.capacity
                and all(self._done(s.started) for s in rung.slots))


# name used by the reference
SHIteration = Bracket


class Hyperband(AbstractPruner):
    def __init__(self, min_budget, max_budget, eta, n_iterations, **kwargs):
        super().__init__(**kwargs)
        if not min_budget > 0:
            raise ValueError(f"Expected `min_budget` > 0, got {min_budget}")
        if min_budget >= max_budget:
            raise ValueError(f"max_budget needs to be larger than min_budget, got {max_budget}, {min_budget}")
        if eta < 2:
            raise ValueError(f"Expected eta greater or equal to 2, got {eta}")
        self.min_budget, self.max_budget, self.eta = min_budget, max_budget, eta
        self.n_iterations = n_iterations
        # number of rungs of the most aggressive bracket and the geometric budget ladder
        self.max_sh_rungs = -int(math.log(min_budget / max_budget) / math.log(eta)) + 1
        self.budgets = [int(max_budget * eta ** (-(self.max_sh_rungs - 1 - i))) for i in range(self.max_sh_rungs)]
        self.iterations: List[Bracket] = []
        self.init_iterations()
        self.start_next_iteration()
        self.updating_iteration = None

    def init_iterations(self):
        s_max = self.max_sh_rungs
        for it in range(self.n_iterations):
            n_rungs = s_max - 1 - (it % s_max)
            n0 = int(math.floor(s_max / (n_rungs + 1)) * self.eta ** n_rungs)
            ns = [max(int(n0 * self.eta ** (-i)), 1) for i in range(n_rungs + 1)]
            self.iterations.append(Bracket(ns, self.budgets[-n_rungs - 1:], it, self.trial_metric_getter, self._log))

    def active_iterations(self):
        return [b for b in self.iterations if b.state == Bracket.RUNNING]

    def start_next_iteration(self):
        for b in self.iterations:
            if b.state == Bracket.INIT:
                b.state = Bracket.RUNNING
                self._log(f"{b.iteration_id}. Iteration started. n_configs: {b.n_configs}, budgets: {b.budgets}")
                self.n_iterations -= 1
                return

    def pruning_routine(self):
        while True:
            for b in self.active_iterations():
                run = b.get_next_run()
                if run is not None:
                    self.updating_iteration = b.iteration_id
                    self._log(f"{b.iteration_id}. Iteration, {b.current_rung}. Rung. Run next {run}")
                    return run
            if self.n_iterations > 0:
                self.start_next_iteration()
                continue
            if self.finished():
                self._log("All Iterations have finished")
                self._close_log()
                return None
            self._log("All Iterations have been started and all trials in their current rung have been started. "
                      "Wait until new run can be scheduled")
            return "IDLE"

    def finished(self) -> bool:
        return all(b.state == Bracket.FINISHED for b in self.iterations)

    def num_trials(self) -> int:
        return sum(sum(b.n_configs) for b in self.iterations)

    def report_trial(self, original_trial_id, new_trial_id):
        self.iterations[self.updating_iteration].report_trial(original_trial_id, new_trial_id)
        self.updating_iteration = None
