"""Driver for distributed PyTorch training (reference:
maggy/core/experiment_driver/torch_distributed_training_driver.py:28-146).

Every GPU gets one worker; all of them run the same train_fn as one data-parallel job. The rendezvous
the reference does over its RPC (REG -> QUERY poll -> EXEC_CONFIG, torch_dist_executor.py:116-130) is
replaced by facts in the assign payload: (rank, world_size, master port). Each worker sends FINAL with
its metric; the experiment result is their mean, ``{"test result": mean}``.
"""
from __future__ import annotations

import json
import os
import time

from ... import util
from ..environment.singleton import EnvSing
from ..workerpool import free_port
from .base import Driver

RUN_ID64 = "00000000000000d1"


class TorchDistributedTrainingDriver(Driver):
    executor_kind = "torch_dist"
    experiment_type = "torch_dist"
    share_devices = True   # one collective job: every rank must see every GPU (symmetric memory / P2P)

    def __init__(self, config, app_id, run_id):
        super().__init__(config, app_id, run_id)
        wanted = getattr(config, "num_workers", None)
        self.num_executors = wanted or util.num_executors(default=2)
        self.results = []
        self._registered = set()
        self._finished = set()
        self._port = free_port()
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
            # a rank of a collective job died: the whole job is lost, fail loudly instead of hanging in NCCL
            self.exception = RuntimeError(f"distributed training worker {w} died and was restarted; "
                                          "collective jobs cannot continue after losing a rank")
            return
        self._registered.add(w)
        payload = {"params": {}, "rank": w, "world_size": self.num_executors, "master_addr": "127.0.0.1",
                   "master_port": self._port}
        while not self.channel.assign(w, RUN_ID64, payload):
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

    def _exp_final_callback(self, job_end: float, exp_json: dict) -> dict:
        self.job_end = job_end
        self.duration = util.seconds_to_milliseconds(job_end - self.job_start)
        result = self.average_metric()
        self.result = {"test result": result}
        env = EnvSing.get_instance()
        env.dump(json.dumps(self.result, default=util.json_default_numpy), os.path.join(self.log_dir, "result.json"))
        env.dump(json.dumps({"name": self.name, "module": "maggy", "app_id": str(self.app_id), "logdir": self.log_dir,
                             "executors": self.num_executors, "description": self.description,
                             "experiment_type": "torch_dist", "status": "FINISHED", "duration": self.duration,
                             "metric": result}, default=util.json_default_numpy),
                 os.path.join(self.log_dir, "maggy.json"))
        util.finalize_experiment(exp_json, result, self.app_id, self.run_id, "FINISHED", self.duration, self.log_dir,
                                 None, None)
        print(f"Final average test loss: {result}" if result is not None else "No metric returned")
        print(f"Finished experiment. Total run time: {util.time_diff(self.job_start, job_end)}")
        return self.result
