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
        self.main_metric_key = None

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

    # ---- the reference's BaseDriver surface (base_driver.py:59-142, 210-254), on top of the in-process run -----------
    def _update_result(self, final_metric) -> None:
        """Fold one run's return value into ``result_dict`` ({key: [values]}); the first key is the main metric."""
        if isinstance(final_metric, dict):
            if self.main_metric_key is None and final_metric:
                self.main_metric_key = next(iter(final_metric))
            for key, value in final_metric.items():
                self.result_dict.setdefault(key, []).append(value)

    def _final_msg_callback(self, msg: dict) -> None:
        final_metric = msg.get("data", None)
        self.results.append(final_metric)
        self._update_result(final_metric)

    def _log_msg_callback(self, msg: dict) -> None:
        if msg.get("data"):
            self.add_executor_logs(msg["data"] if isinstance(msg["data"], str) else str(msg["data"]))

    def average_metrics(self) -> float:
        """Mean of the numeric results so far (0 if there is none yet)."""
        vals = [r for r in self.results if isinstance(r, (int, float)) and not isinstance(r, bool)]
        return sum(vals) / len(vals) if vals else 0

    def prep_results(self, duration_str: str) -> str:
        return f"Results ------\nMetrics value {self.result}\nTotal job time {duration_str}\n"

    def json(self) -> str:
        exp = {"project": EnvSing.get_instance().project_name(), "user": os.environ.get("HADOOP_USER_NAME"),
               "name": self.name, "module": "maggy", "app_id": str(self.app_id),
               "start": __import__("time").strftime("%Y-%m-%dT%H:%M:%S", __import__("time").localtime(self.job_start)),
               "executors": self.num_executors, "logdir": self.log_dir, "description": self.description,
               "experiment_type": "base"}
        if self.experiment_done:
            exp.update({"status": "FINISHED", "duration": self.duration,
                        "finished": __import__("time").strftime("%Y-%m-%dT%H:%M:%S",
                                                                __import__("time").localtime(self.job_end))})
            if self.main_metric_key is not None and self.result_dict.get(self.main_metric_key):
                exp["metric"] = self.result_dict[self.main_metric_key][-1]
        else:
            exp["status"] = "RUNNING"
        return json.dumps(exp, default=util.json_default_numpy)

    def finalize(self, job_end: float) -> dict:
        """Summarise the run: ``result.json`` + ``maggy.json`` in the log directory, the summary line in ``maggy.log``."""
        self.job_end = job_end
        self.duration = util.seconds_to_milliseconds(job_end - self.job_start)
        for r in self.results:
            self._update_result(r)
        self.result = {"test result": self.average_metric()}
        summary = self.prep_results(util.time_diff(self.job_start, job_end))
        self.log(summary)
        env = EnvSing.get_instance()
        env.dump(json.dumps(self.result, default=util.json_default_numpy), os.path.join(self.log_dir, "result.json"))
        env.dump(self.json(), os.path.join(self.log_dir, "maggy.json"))
        return self.result

    def _exp_final_callback(self, job_end: float, exp_json: dict) -> dict:
        result = self.finalize(job_end)
        util.finalize_experiment(exp_json, None, self.app_id, self.run_id, "FINISHED", self.duration, self.log_dir,
                                 None, None)
        print(f"List of the results:\n{self.result}" if self.results else "No results")
        return result

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
