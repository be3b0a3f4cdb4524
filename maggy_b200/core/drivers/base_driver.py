"""Driver for plain single runs: ``lagom(train_fn)`` / ``lagom(train_fn, BaseConfig())``
(reference: maggy/core/experiment_driver/base_driver.py:35-258 and python_driver.py:135-143).

train_fn runs once in the calling process — no worker pool, no control segment; a Reporter is still
handed in so ``reporter.broadcast`` / ``reporter.log`` work. The result is ``{"test result": ...}``.
"""
from __future__ import annotations

import json
import os

from ... import util
from ..environment.singleton import EnvSing
from ..executors.base_executor import base_executor_fn
from ..reporter import Reporter
from .base import Driver


class BaseDriver(Driver):
    in_process = True
    executor_kind = "base"

    def __init__(self, config, app_id, run_id):
        super().__init__(config, app_id, run_id)
        self.num_executors = 1
        self.results = []
        self.result_dict = {}

    def _exp_startup_callback(self) -> None:
        pass

    def _register_msg_callbacks(self) -> None:
        pass

    def _run_in_process(self, train_fn, config):
        reporter = Reporter(os.path.join(self.log_dir, "executor_0_0.log"), 0, 0, None, None)
        reporter.init_logger(os.path.join(self.log_dir, "output.log"))
        try:
            retval = base_executor_fn(train_fn, config, reporter)()
        finally:
            reporter.close_logger()
        self.results.append(retval)
        self.experiment_done = True

    def _exp_final_callback(self, job_end: float, exp_json: dict) -> dict:
        self.job_end = job_end
        self.duration = util.seconds_to_milliseconds(job_end - self.job_start)
        result = self.average_metric()
        self.result = {"test result": result}
        env = EnvSing.get_instance()
        env.dump(json.dumps(self.result, default=util.json_default_numpy), os.path.join(self.log_dir, "result.json"))
        env.dump(json.dumps({"name": self.name, "module": "maggy", "app_id": str(self.app_id), "logdir": self.log_dir,
                             "description": self.description, "experiment_type": "base", "status": "FINISHED",
                             "duration": self.duration}, default=util.json_default_numpy),
                 os.path.join(self.log_dir, "maggy.json"))
        util.finalize_experiment(exp_json, None, self.app_id, self.run_id, "FINISHED", self.duration, self.log_dir,
                                 None, None)
        print(f"List of the results:\n{self.result}" if self.results else "No results")
        return self.result

    def average_metric(self):
        """dict results -> {key: [values over runs]}; numbers -> their mean; anything else as is."""
        vals = [r for r in self.results if r is not None]
        if not vals:
            return None
        if all(isinstance(v, dict) for v in vals):
            out = {}
            for v in vals:
                for k, x in v.items():
                    out.setdefault(k, []).append(x)
            return out
        if all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in vals):
            return sum(vals) / len(vals)
        return vals[0] if len(vals) == 1 else vals
