"""Driver for distributed TensorFlow training (reference:
maggy/core/experiment_driver/tf_distributed_training_driver.py:37-271 + ``TensorflowServer``, maggy/core/rpc.py:593-633).

Every executor registers; once all of them have, the driver derives the cluster specification from the
registrations (``MirroredReservations``: first = chief, last = evaluator, the rest workers — the reference relabels
the last worker as evaluator in the executor, tf_dist_executor.py:123-148) and hands each task its own ``TF_CONFIG``
inside the assignment. That replaces the reference's three poll loops (QUERY until all registered, RESERVATIONS,
TF_CONFIG) with one write per worker. FINAL metrics are averaged: ``{"test result": mean}``.
"""
from __future__ import annotations

import json
import os
import time

from ... import util
from ..environment.singleton import EnvSing
from ..reservations import MirroredReservations
from ..workerpool import free_port
from .base import Driver

RUN_ID64 = "00000000000000f1"


def tf_config_for(spec: dict, host_port: str) -> dict:
    """TF_CONFIG of the task that listens on ``host_port`` inside cluster ``spec`` (roles -> [host:port, ...])."""
    for role, hosts in spec.items():
        if host_port in hosts:
            return {"cluster": {r: list(h) for r, h in spec.items()}, "task": {"type": role, "index": hosts.index(host_port)}}
    raise KeyError(f"{host_port} is not part of the cluster {spec}")


class TfDistributedTrainingDriver(Driver):
    executor_kind = "tf_dist"
    experiment_type = "tf_dist"
    share_devices = True

    def __init__(self, config, app_id, run_id):
        super().__init__(config, app_id, run_id)
        wanted = getattr(config, "num_workers", None)
        self.num_executors = wanted or util.num_executors(default=2)
        self.mirrored = MirroredReservations(self.num_executors, evaluator=self.num_executors >= 2)
        self.results = []
        self._finished = set()
        self.cluster_spec = None
        self.result = {}

    def _exp_startup_callback(self) -> None:
        pass

    def _register_msg_callbacks(self) -> None:
        self.message_callbacks.update({
            "REG": self._register_msg_callback, "FINAL": self._final_msg_callback,
            "METRIC": lambda msg: None, "LOG": self._log_msg_callback, "ERROR": self._error_msg_callback,
            "EXIT": lambda msg: None, "IDLE": lambda msg: None,
        })

    def _register_msg_callback(self, msg):
        w = msg["partition_id"]
        if msg["aux"] > 0:
            self.exception = RuntimeError(f"distributed training worker {w} died and was restarted; "
                                          "a MultiWorkerMirroredStrategy job cannot continue after losing a task")
            return
        # single host: the driver hands out the listen address of every task
        self.mirrored.add({"partition_id": w, "task_attempt": msg["aux"], "host_port": f"127.0.0.1:{free_port()}",
                           "trial_id": None})
        if not self.mirrored.done():
            return
        self.cluster_spec = self.mirrored.get_spec()
        self.log(f"TensorFlow cluster: {self.cluster_spec}")
        training = self.cluster_spec.get("chief", []) + self.cluster_spec.get("worker", [])
        for pid, r in sorted(self.mirrored.get().items()):
            cfg = tf_config_for(self.cluster_spec, r["host_port"])
            payload = {"params": {}, "tf_config": cfg, "world_size": self.num_executors,
                       "num_training_tasks": max(1, len(training)),
                       "training_index": training.index(r["host_port"]) if r["host_port"] in training else 0}
            while not self.channel.assign(pid, RUN_ID64, payload):
                time.sleep(0.001)

    def _final_msg_callback(self, msg):
        data = msg["data"] or {}
        self.results.append(data.get("metric", msg["value"]))
        self._finished.add(msg["partition_id"])
        if data.get("logs"):
            self.add_executor_logs(data["logs"])
        if len(self._finished) >= self.num_executors:
            self.experiment_done = True

    def _log_msg_callback(self, msg):
        if msg["data"]:
            self.add_executor_logs(msg["data"])

    def _error_msg_callback(self, msg):
        text = (msg["data"] or {}).get("error", "") if isinstance(msg["data"], dict) else str(msg["data"])
        self.log(f"worker {msg['partition_id']} failed:\n{text}")
        self.exception = RuntimeError(f"train_fn raised on worker {msg['partition_id']}:\n{text}")

    def average_metric(self):
        vals = [v for v in self.results if isinstance(v, (int, float))]
        return sum(vals) / len(vals) if vals else None

    # ---- the reference driver's result surface (tf_distributed_training_driver.py:60-142, 262-271) ------------------
    def set_result(self, result) -> None:
        self.results.append(result)

    def set_final_model(self, final_model) -> None:
        self.final_model = final_model

    def _update_result(self, final_metric) -> None:
        self.set_result(final_metric)

    def prep_results(self, duration_str: str) -> str:
        return f"Results ------\nAverage metric over {self.num_executors} workers {self.average_metric()}\n" \
               f"Total job time {duration_str}\n"

    def json(self) -> str:
        exp = {"name": self.name, "module": "maggy", "app_id": str(self.app_id), "logdir": self.log_dir,
               "executors": self.num_executors, "description": self.description, "experiment_type": "tf_dist",
               "cluster": self.cluster_spec}
        if self.experiment_done:
            exp.update({"status": "FINISHED", "duration": self.duration, "metric": self.average_metric()})
        else:
            exp["status"] = "RUNNING"
        return json.dumps(exp, default=util.json_default_numpy)

    def finalize(self, job_end: float) -> dict:
        self.job_end = job_end
        self.duration = util.seconds_to_milliseconds(job_end - self.job_start)
        self.result = {"test result": self.average_metric()}
        self.log(self.prep_results(util.time_diff(self.job_start, job_end)))
        env = EnvSing.get_instance()
        env.dump(json.dumps(self.result, default=util.json_default_numpy), os.path.join(self.log_dir, "result.json"))
        env.dump(self.json(), os.path.join(self.log_dir, "maggy.json"))
        return self.result

    def _exp_final_callback(self, job_end: float, exp_json: dict) -> dict:
        result = self.finalize(job_end)["test result"]
        util.finalize_experiment(exp_json, result, self.app_id, self.run_id, "FINISHED", self.duration, self.log_dir,
                                 None, None)
        print(f"Final average test loss: {result}" if result is not None else "No metric returned")
        print(f"Finished experiment. Total run time: {util.time_diff(self.job_start, job_end)}")
        return self.result
