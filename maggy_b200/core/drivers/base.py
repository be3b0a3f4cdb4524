"""Experiment driver skeleton (reference: maggy/core/experiment_driver/spark_driver.py:39-286 and
python_driver.py:39-267 — here ONE class, no Spark/Python fork).

Template method ``run_experiment``: startup hook -> control segment + message digest thread -> worker
pool (native launcher, one process per GPU / GPU gang) -> supervise until the experiment is done ->
final hook (aggregate result, write artefacts) -> stop. Subclasses register message callbacks
(``REG/METRIC/FINAL/IDLE/LOG/ERROR/EXIT``) and implement the hooks.

Differences to the reference that matter for throughput: no 1 s server ``select`` timeout, no busy-spin
digest loop (``_digest_queue`` in the reference burns a core, spark_driver.py:219-236) — the digest
thread blocks in the native pump's condition variable — and no fixed 6 s grace sleep at the end.
"""
from __future__ import annotations

import os
import threading
import time
import traceback
from abc import ABC, abstractmethod
from datetime import datetime
from typing import Callable, Dict, Optional

import cloudpickle

from ... import util
from .. import rpc
from ..environment.singleton import EnvSing
from ..exceptions import WorkerLostError
from ..workerpool import WorkerPool


class Driver(ABC):
    """Owns the control channel, the worker pool, the message digest thread and ``maggy.log``."""

    executor_kind = "trial"   # which loop the workers run (see core/executors/worker_main.py)
    in_process = False        # True: run train_fn in the driver process, no pool (BaseDriver)

    def __init__(self, config, app_id, run_id):
        self.config = config
        self.app_id = app_id
        self.run_id = run_id
        self.name = config.name
        self.description = config.description
        self.hb_interval = config.hb_interval
        self.result = None
        self.job_start = None
        self.job_end = None
        self.duration = None
        self.experiment_done = False
        self.worker_done = False
        self.exception = None
        self.channel: Optional[rpc.DriverChannel] = None
        self.pool: Optional[WorkerPool] = None
        self.num_executors = 1
        self.gpus_per_worker = 1
        self.message_callbacks: Dict[str, Callable] = {}
        self.executor_logs = ""
        self.log_lock = threading.RLock()
        env = EnvSing.get_instance()
        self.log_dir = env.get_logdir(app_id, run_id)
        self.log_file = os.path.join(self.log_dir, "maggy.log")
        if not env.exists(self.log_file):
            env.dump("", self.log_file)
        self.fd = env.open_file(self.log_file, flags="a")
        self._digest_thread = None
        self._respawns = 0
        self.max_respawns = int(os.environ.get("MAGGY_MAX_RESPAWNS", "8"))
        self._gstop_sent = False

    # ------------------------------------------------------------------ hooks
    @abstractmethod
    def _exp_startup_callback(self) -> None:
        ...

    @abstractmethod
    def _exp_final_callback(self, job_end: float, exp_json: dict) -> dict:
        ...

    def _exp_exception_callback(self, exc: Exception) -> None:
        self.log(f"Experiment failed: {exc!r}")

    @abstractmethod
    def _register_msg_callbacks(self) -> None:
        ...

    def _worker_spec(self, train_fn: Callable, config) -> dict:
        """What every worker process needs (the analogue of the closure Spark ships to executors)."""
        env = EnvSing.get_instance()
        return {
            "kind": self.executor_kind, "train_fn": train_fn, "config": self._picklable_config(config),
            "log_dir": self.log_dir, "env_log_dir": env.log_dir, "app_id": self.app_id, "run_id": self.run_id,
            "hb_interval": self.hb_interval, "experiment_type": getattr(self, "experiment_type", "optimization"),
            "optimization_key": getattr(config, "optimization_key", "Metric"),
            "sys_path": [p for p in __import__("sys").path if p],
        }

    def _picklable_config(self, config):
        return config

    def _patching_fn(self, train_fn: Callable, config, is_spark_available: bool = False) -> Callable:
        """``executor_fn(partition_id)`` for this driver's executor kind (reference: every driver's ``_patching_fn``
        returns the closure Spark runs on its executors). The native pool normally starts the same loop through
        ``worker_main``; calling the returned function runs it in the current process against this driver's segment."""
        from ..executors.worker_main import run_executor

        spec, segment = self._worker_spec(train_fn, config), self.channel.name

        def executor_fn(partition_id, attempt: int = 0):
            return run_executor(spec, segment, int(partition_id), attempt)

        return executor_fn

    def _on_tick(self) -> None:
        """Called every digest iteration (idle retries, timers)."""

    # ------------------------------------------------------------------ template method
    def run_experiment(self, train_fn: Callable, config) -> dict:
        job_start = time.time()
        self.job_start = job_start
        try:
            self._exp_startup_callback()
            exp_json = util.populate_experiment(config, self.app_id, self.run_id, train_fn.__name__)
            self.log(f"Started Maggy Experiment: {self.name}, {self.app_id}, run {self.run_id}")
            if self.in_process:
                self._run_in_process(train_fn, config)
            else:
                self.init(job_start)
                self._launch_workers(train_fn, config)
                self._supervise()
            job_end = time.time()
            result = self._exp_final_callback(job_end, exp_json)
            return result
        except Exception as exc:  # noqa: BLE001
            self._exp_exception_callback(exc)
            raise
        finally:
            self.stop()

    def _run_in_process(self, train_fn, config):
        raise NotImplementedError

    def init(self, job_start: float) -> None:
        """Create the control segment and start the digest thread (reference: server.start + _start_worker)."""
        self.job_start = job_start
        self._register_msg_callbacks()
        self.channel = rpc.DriverChannel(self.num_executors, ring_slots=1024, payload_bytes=4 << 20)
        self._start_worker()
        self._start_log_server()

    def _start_log_server(self) -> None:
        """External LOG endpoint (reference rpc.py:490-502): on request (MAGGY_LOG_SERVER=1 | host:port) or whenever
        the environment wants to announce the driver (Hopsworks registers it over REST for its Jupyter front-end)."""
        self.log_server = None
        want = os.environ.get("MAGGY_LOG_SERVER", "")
        env = EnvSing.get_instance()
        announce = getattr(env, "announce_driver", None)
        if not want and announce is None:
            return
        try:
            from ..logserver import LogServer

            host, port = "127.0.0.1", 0
            if ":" in want:
                host, port = want.rsplit(":", 1)[0], int(want.rsplit(":", 1)[1])
            elif announce is not None:
                host = ""
            self.log_server = LogServer(self, host, port, secret=getattr(self, "_secret", None)).start()
            self._secret = self.log_server.secret
            self.log(f"LOG endpoint listening on {self.log_server.address[0] or '0.0.0.0'}:{self.log_server.address[1]} "
                     f"secret {self._secret}")
            if announce is not None:
                announce(self.log_server.address[1], self._secret, self.app_id, self)
        except Exception as e:  # noqa: BLE001 - observability must never fail the experiment
            self.log(f"LOG endpoint not started: {e}")

    def _scheduler_gpus(self):
        """Physical GPU(s) the workers must be able to map besides their own (the device scheduler's mailbox)."""
        sched = getattr(self, "device_scheduler", None)
        if sched is not None and sched.mailbox:
            return [int(sched.device.index or 0)]
        return []

    def _launch_workers(self, train_fn, config):
        spec = self._worker_spec(train_fn, config)
        spec_path = os.path.join(self.log_dir, ".worker_spec.pkl")
        with open(spec_path, "wb") as f:
            cloudpickle.dump(spec, f)
        self.pool = WorkerPool(self.num_executors, self.channel.name, spec_path, self.log_dir,
                               gpus_per_worker=self.gpus_per_worker,
                               share_devices=getattr(self, "share_devices", False),
                               also_visible=self._scheduler_gpus())
        self.pool.start()

    # ------------------------------------------------------------------ supervision (main thread)
    def _supervise(self):
        """Block until the experiment is done; respawn dead workers (reference: Spark task retry +
        re-REG => BLACK, rpc.py:421-432)."""
        liveness_timeout = max(10.0 * float(self.hb_interval), 30.0)
        done_since = None
        while True:
            if self.exception is not None:
                self._send_gstop()
                raise self.exception
            for w in range(self.num_executors):
                state = self.channel.worker_state(w)
                if self.pool.alive(w):
                    stuck = (state in (rpc.W_RUNNING, rpc.W_IDLE, rpc.W_REGISTERED)
                             and self.channel.heartbeat_age_s(w) > liveness_timeout)
                    if stuck and not self.experiment_done:
                        self.log(f"worker {w} has not shown a sign of life for {liveness_timeout:.0f}s: restarting it")
                        self._respawn(w)
                    continue
                if state == rpc.W_DONE or self.experiment_done:
                    continue
                info = self.pool.exit_info(w)
                self.log(f"worker {w} died unexpectedly {info}: restarting it")
                self._respawn(w)
            if self.experiment_done:
                self._send_gstop()
                if done_since is None:
                    done_since = time.time()
                if all(not self.pool.alive(w) for w in range(self.num_executors)) or time.time() - done_since > 10:
                    break
            time.sleep(0.02)
        self.pool.shutdown(grace_s=2.0)
        # drain what the workers posted on their way out
        time.sleep(0.02)
        self.worker_done = True
        if self._digest_thread is not None:
            self._digest_thread.join(timeout=5)
        if self.exception is not None:
            raise self.exception

    def _respawn(self, worker: int):
        self._respawns += 1
        if self._respawns > self.max_respawns:
            raise WorkerLostError(f"worker {worker} keeps dying ({self._respawns} restarts); see "
                                  f"{self.log_dir}/worker_{worker}_0.stdout.log")
        self.pool.respawn(worker)

    def _send_gstop(self):
        if not self._gstop_sent and self.channel is not None:
            self.channel.gstop()
            self._gstop_sent = True

    # ------------------------------------------------------------------ digest thread
    def add_message(self, msg: dict) -> None:
        """Hand one message to its callback (reference: spark_driver.py add_message puts it on the queue the digest
        thread drains; here records arrive through the control segment, this entry serves injected messages — tests,
        the external LOG endpoint)."""
        cb = self.message_callbacks.get(msg["type"])
        if cb is not None:
            cb(msg)

    @staticmethod
    def _generate_secret(nbytes: int = 8) -> str:
        """Random hex token (reference: spark_driver.py:70-73 authenticates TCP clients with it). The control segment's
        name carries such a token and the segment is created 0600: knowing the name IS the credential."""
        import secrets

        return secrets.token_hex(nbytes)

    def _start_worker(self) -> None:
        """Start the digest thread (reference: spark_driver.py:211-217)."""
        self._digest_thread = threading.Thread(target=self._digest_loop, name="maggy-digest", daemon=True)
        self._digest_thread.start()

    def _digest_loop(self):
        try:
            while not self.worker_done:
                ch = self.channel
                if ch is None:
                    break
                for msg in ch.next(timeout_s=0.02):
                    self.add_message(msg)
                self._on_tick()
        except Exception as exc:  # noqa: BLE001
            self.log(traceback.format_exc())
            self.exception = exc

    # ------------------------------------------------------------------ logging
    def log(self, log_msg) -> None:
        with self.log_lock:
            try:
                self.fd.write(EnvSing.get_instance().str_or_byte(f"{datetime.now().isoformat()}: {log_msg}\n"))
                self.fd.flush()
            except ValueError:
                pass

    def add_executor_logs(self, text: str):
        with self.log_lock:
            self.executor_logs += text

    def get_logs(self):
        """(progress/result so far, drained executor logs) — what the reference's LOG message returns."""
        with self.log_lock:
            logs, self.executor_logs = self.executor_logs, ""
            return self.result, logs

    def stop(self) -> None:
        self.worker_done = True
        try:
            from ...utils import tracing

            tracing.stop()
        except Exception:  # noqa: BLE001
            pass
        try:
            if self.pool is not None:
                self.pool.shutdown(grace_s=1.0)
        finally:
            self.pool = None
        if getattr(self, "log_server", None) is not None:
            self.log_server.stop()
            self.log_server = None
        if self.channel is not None:
            try:
                self._send_gstop()
            except Exception:
                pass
            if self._digest_thread is not None and self._digest_thread.is_alive():
                self._digest_thread.join(timeout=2)
            self.channel.close()
            self.channel = None
        with self.log_lock:
            if not self.fd.closed:
                self.fd.flush()
                self.fd.close()
