"""Worker-side reporter handed to the user's train_fn (reference: maggy/core/reporter.py:30-170).

Same contract — ``broadcast(metric, step=None)`` validates the metric/step, records them, and raises
``EarlyStopException`` when the trial has been told to stop; ``log(msg)`` tees into the executor log,
the trial log and the driver — but no heartbeat thread ships the data once per second:

* a numeric metric becomes a 64-byte ring record immediately (sub-millisecond to the driver);
* a 0-dim CUDA tensor is published by a one-thread kernel on the current stream (K8), so the training
  loop never synchronises on ``loss.item()``;
* the stop decision is one atomic load of the worker's stop word, so a trial reacts at the very next
  ``broadcast`` instead of after one to two heartbeat periods.
"""
from __future__ import annotations

import threading
import time
from datetime import datetime

from .. import constants
from . import exceptions
from .environment.singleton import EnvSing


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "dim")


class Reporter:
    def __init__(self, log_file, partition_id, task_attempt, print_executor=None, channel=None):
        self.metric = None
        self.step = -1
        self.lock = threading.RLock()
        self.stop = False
        self.trial_id = None
        self.trial_log_file = None
        self.logs = ""
        self.log_file = log_file
        self.partition_id = partition_id
        self.task_attempt = task_attempt
        self.print_executor = print_executor or (lambda *a, **k: None)
        self.channel = channel
        self._last_log_flush = 0.0
        # proxy of the reference's mechanism (benchmarks only): metrics are not sent by broadcast() but by a
        # heartbeat thread calling ship() once per hb_interval, and STOP arrives with that heartbeat
        self.deferred = False
        self._shipped_step = None
        # multi-GPU trial gangs: only the leader hears STOP from the driver, but every member must leave
        # train_fn at the same broadcast() call or the others hang in their next collective
        self.gang = None
        env = EnvSing.get_instance()
        if log_file:
            if not env.exists(log_file):
                env.dump("", log_file)
            self.fd = env.open_file(log_file, flags="a")
        else:
            self.fd = None
        self.trial_fd = None

    # ------------------------------------------------------------------ trial log files
    def init_logger(self, trial_log_file):
        self.trial_log_file = trial_log_file
        env = EnvSing.get_instance()
        if not env.exists(trial_log_file):
            env.dump("", trial_log_file)
        self.trial_fd = env.open_file(trial_log_file, flags="a")

    def close_logger(self):
        with self.lock:
            if self.trial_fd:
                self.trial_fd.close()
                self.trial_fd = None
            if self.fd:
                self.fd.close()
                self.fd = None

    # ------------------------------------------------------------------ metrics
    def broadcast(self, metric, step=None):
        """Report ``metric`` for ``step`` (default: previous step + 1) to the experiment driver.

        Raises ``EarlyStopException`` if the driver / policy engine stopped this trial."""
        with self.lock:
            if step is None:
                step = self.step + 1
            on_device = False
            if _is_tensor(metric):
                if metric.numel() != 1:
                    raise exceptions.BroadcastMetricTypeError(metric)
                on_device = bool(metric.is_cuda) and self.channel is not None
                if not on_device:
                    metric = metric.item()
            elif isinstance(metric, bool) or not isinstance(metric, constants.USER_FCT.NUMERIC_TYPES):
                raise exceptions.BroadcastMetricTypeError(metric)
            if isinstance(step, bool) or not isinstance(step, constants.USER_FCT.NUMERIC_TYPES):
                raise exceptions.BroadcastStepTypeError(metric, step)
            if step < self.step:
                raise exceptions.BroadcastStepValueError(metric, step, self.step)
            self.step = step
            self.metric = metric
            if self.channel is not None and self.trial_id is not None and not self.deferred:
                if on_device:
                    self.channel.post_metric_device(self.trial_id, metric.detach().reshape(()), int(step))
                else:
                    self.channel.post_metric(self.trial_id, float(metric), int(step))
                if not self.stop and self.channel.check_stop(self.trial_id):
                    self.stop = True
                self._maybe_flush_logs()
            if self.gang is not None:
                # agreement is reached one call later: all members raise at the same call index
                if self.gang.agree_stop(self.stop):
                    self.stop = True
                    raise exceptions.EarlyStopException(self._metric_value())
            elif self.stop:
                raise exceptions.EarlyStopException(self._metric_value())

    def ship(self):
        """Reference-style heartbeat: send the latest metric (if new) and pick up a pending STOP."""
        with self.lock:
            if self.channel is None or self.trial_id is None or self.metric is None:
                return
            if self._shipped_step != self.step:
                self.channel.post_metric(self.trial_id, float(self._metric_value()), int(self.step))
                self._shipped_step = self.step
            if self.channel.check_stop(self.trial_id):
                self.early_stop()

    def _metric_value(self):
        m = self.metric
        return m.item() if _is_tensor(m) else m

    def early_stop(self):
        """Mark the running trial as stopped (only meaningful once a metric has been reported)."""
        with self.lock:
            if self.metric is not None:
                self.stop = True

    def get_data(self):
        """(metric, step, drained logs) — what the reference's heartbeat would have shipped."""
        with self.lock:
            logs, self.logs = self.logs, ""
            return self._metric_value(), self.step, logs

    # ------------------------------------------------------------------ logs
    def log(self, log_msg, jupyter=False):
        """Write to the executor log + trial log (and buffer for the driver when ``jupyter=True``)."""
        with self.lock:
            env = EnvSing.get_instance()
            try:
                msg = f"{datetime.now().isoformat()} ({self.partition_id}/{self.task_attempt}): {log_msg} \n"
                if jupyter:
                    if self.trial_fd:
                        self.trial_fd.write(env.str_or_byte(msg))
                    self.logs += f"{self.partition_id}: {log_msg}\n"
                else:
                    if self.fd:
                        self.fd.write(env.str_or_byte(msg))
                    if self.trial_fd:
                        self.trial_fd.write(env.str_or_byte(msg))
                    self.print_executor(msg)
            except (IOError, ValueError, AttributeError) as e:
                if self.fd:
                    self.fd.write(env.str_or_byte(f"An error occurred while writing logs: {e}"))

    def _maybe_flush_logs(self, force=False):
        if self.channel is None or not self.logs:
            return
        now = time.monotonic()
        if force or now - self._last_log_flush > 0.5:
            logs, self.logs = self.logs, ""
            self._last_log_flush = now
            self.channel.post_log(self.trial_id, logs)

    # ------------------------------------------------------------------ per-trial lifecycle
    def reset(self):
        with self.lock:
            self._maybe_flush_logs(force=True)
            self.metric = None
            self.step = -1
            self.stop = False
            self.trial_id = None
            self._shipped_step = None
            if self.fd:
                self.fd.flush()
            if self.trial_fd:
                self.trial_fd.close()
            self.trial_fd = None
            self.trial_log_file = None

    def get_trial_id(self):
        with self.lock:
            return self.trial_id

    def set_trial_id(self, trial_id):
        with self.lock:
            self.trial_id = trial_id
