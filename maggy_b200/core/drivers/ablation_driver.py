"""Ablation-study driver (reference: maggy/core/experiment_driver/ablation_driver.py:32-208).

Same machinery as hyperparameter optimization with an ablator as the controller and no early
stopping. (The reference class is broken as written — wrong ``_patching_fn`` arity and a missing
``OptimizationServer`` argument, SURVEY.md 2.8 — this one runs.)
"""
from __future__ import annotations

import json
import threading

from ... import util
from ...ablation.ablationstudy import AblationStudy
from ...ablation.ablator import LOCO, AbstractAblator
from ...earlystop import NoStoppingRule
from .optimization_driver import HyperparameterOptDriver


class AblationDriver(HyperparameterOptDriver):
    experiment_type = "ablation"

    def __init__(self, config, app_id, run_id):
        # bypass HyperparameterOptDriver.__init__ (no searchspace/optimizer) but reuse Driver.__init__
        from .base import Driver

        Driver.__init__(self, config, app_id, run_id)
        self._final_store = []
        self._trial_store = {}
        self._store_lock = threading.RLock()
        self.maggy_log = ""
        self.searchspace = None
        self.optimization_key = getattr(config, "optimization_key", "Metric")
        self.direction = self._init_direction(config.direction)
        self.es_policy = NoStoppingRule
        self.earlystop_check = NoStoppingRule.earlystop_check
        self._no_earlystop = True
        self.es_interval, self.es_min = 1, 10 ** 9
        if not isinstance(config.ablation_study, AblationStudy):
            raise Exception(f"The experiment's ablation study configuration should be an instance of "
                            f"maggy.ablation.AblationStudy, but it is {config.ablation_study} (of type "
                            f"'{type(config.ablation_study).__name__}').")
        self.ablation_study = config.ablation_study
        ablator = config.ablator
        if isinstance(ablator, str):
            if ablator.lower() != "loco":
                raise Exception(f"The experiment's ablation study policy should either be a string ('loco') or a "
                                f"custom policy that is an instance of maggy.ablation.ablation.AbstractAblator, but "
                                f"it is {ablator} (of type '{type(ablator).__name__}').")
            self.controller = LOCO(config.ablation_study, self._final_store)
        elif isinstance(ablator, AbstractAblator):
            self.controller = ablator
            print(f"Custom Ablator initialized: {type(ablator).__name__}")
        else:
            raise Exception(f"The experiment's ablation study policy should either be a string ('loco') or a custom "
                            f"policy that is an instance of maggy.ablation.ablation.AbstractAblator, but it is "
                            f"{ablator} (of type '{type(ablator).__name__}').")
        self.num_trials = self.controller.get_number_of_trials()
        self.gpus_per_worker = max(1, int(getattr(config, "gpus_per_trial", 1) or 1))
        wanted = getattr(config, "num_workers", None)
        pool = wanted or max(1, util.num_executors(default=2) // self.gpus_per_worker)
        self.num_executors = max(1, min(pool, self.num_trials))
        self.result = {"best_val": "n.a.", "num_trials": 0, "early_stopped": "n.a"}
        self.controller.pruner = None
        self.controller.initialize()
        self._assigned, self._idle, self._scheduled = {}, {}, 0
        self._controller_exhausted = False
        self._lookahead_enabled, self._lookahead, self._lookahead_thread = False, None, None
        self._controller_lock = threading.RLock()
        self.inter_trial_gaps, self.stop_latencies = [], []

    def _exp_startup_callback(self) -> None:
        pass

    def controller_get_next(self, trial=None):
        return self.controller.get_trial(trial)

    def init(self, job_start: float) -> None:
        from .base import Driver

        Driver.init(self, job_start)
        self._native_policy = False
        self.channel.configure_policy(False, self.direction, 10 ** 9, 1)

    def prep_results(self, duration_str: str) -> str:
        self.controller.finalize_experiment(self._final_store)
        r = self.result
        return ("\n------ " + self.controller.name() + " Results ------ \n"
                "BEST Config Excludes " + json.dumps(r.get("best_config"), default=util.json_default_numpy)
                + " -- metric " + str(r.get("best_val")) + "\n"
                "WORST Config Excludes " + json.dumps(r.get("worst_config"), default=util.json_default_numpy)
                + " -- metric " + str(r.get("worst_val")) + "\n"
                "AVERAGE metric -- " + str(r.get("avg")) + "\n"
                "Total Job Time " + duration_str + "\n")

    def config_to_dict(self) -> dict:
        return self.ablation_study.to_dict()

    def log_string(self) -> str:
        r = self.result
        return ("Maggy Ablation " + str(r["num_trials"]) + "/" + str(self.num_trials) + " "
                + util.progress_bar(r["num_trials"], self.num_trials)
                + " - BEST Excludes " + json.dumps(r.get("best_config"), default=util.json_default_numpy)
                + " - metric " + str(r["best_val"]))
