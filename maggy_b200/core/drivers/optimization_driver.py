"""Hyperparameter-optimization driver (reference: maggy/core/experiment_driver/optimization_driver.py:40-691).

Keeps the reference's bookkeeping (trial stores, result aggregation, ``trial.json`` / ``result.json`` /
``maggy.json`` artefacts, progress lines) and message semantics (REG / METRIC / FINAL / IDLE / BLACK), on
top of the ring control plane:

* a finishing worker gets its next trial in the same digest turn as its FINAL (no 1 s GET poll);
* the median stopping rule is armed in the native policy engine, which writes the STOP word itself;
  the Python driver only mirrors the decision into the Trial (custom ``AbstractEarlyStop`` policies
  still run here and stop trials through ``channel.stop_trial``);
* model-based controllers get a one-deep lookahead: the next suggestion is computed on a helper thread
  while trials run, so a GP/TPE fit never sits between FINAL and the next assignment.
"""
from __future__ import annotations

import json
import os
import threading
import time
from typing import Optional

from ... import tensorboard, util
from ...earlystop import AbstractEarlyStop, MedianStoppingRule, NoStoppingRule
from ...optimizer import AbstractOptimizer, Asha, GridSearch, RandomSearch, SingleRun
from ...optimizer.bayes import GP, TPE
from ...searchspace import Searchspace
from ...trial import Trial
from ...utils import tracing
from ..environment.singleton import EnvSing
from .base import Driver


class HyperparameterOptDriver(Driver):
    controller_dict = {
        "randomsearch": RandomSearch,
        "asha": Asha,
        "tpe": TPE,
        "gp": GP,
        "none": SingleRun,
        "faulty_none": None,
        "gridsearch": GridSearch,
    }
    executor_kind = "trial"
    experiment_type = "optimization"

    _resumable: dict = {}   # class-level defaults: subclasses with their own constructor (ablation) never resume
    _device_asha = False    # ... and never hand their next-trial decision to the device scheduler
    _staged: dict = {}
    resumed_trials = 0

    def __init__(self, config, app_id, run_id):
        super().__init__(config, app_id, run_id)
        self._final_store = []
        self._trial_store = {}
        self._store_lock = threading.RLock()
        self.experiment_done = False
        self.maggy_log = ""
        self.job_end = None
        self.duration = None
        self.searchspace = self._init_searchspace(config.searchspace)
        self.controller = self._init_controller(config.optimizer, self.searchspace)
        self.optimization_key = getattr(config, "optimization_key", "Metric")
        self.direction = self._init_direction(config.direction)
        self.num_trials = config.num_trials
        if self.controller.pruner:
            self.num_trials = self.controller.pruner.num_trials()
        if isinstance(self.controller, GridSearch):
            self.num_trials = self.controller.get_num_trials(self.searchspace)
        self.gpus_per_worker = max(1, int(getattr(config, "gpus_per_trial", 1) or 1))
        wanted = getattr(config, "num_workers", None)
        pool = wanted or max(1, util.num_executors(default=2) // self.gpus_per_worker)
        self.num_executors = max(1, min(pool, self.num_trials))
        self.es_policy = self._init_earlystop_check(config.es_policy)
        # the bound check itself (reference: optimization_driver.py:657-691 returns the policy's static method)
        self.earlystop_check = self.es_policy.earlystop_check
        self._no_earlystop = self._policy_is(self.es_policy, NoStoppingRule)
        self.es_interval = config.es_interval
        self.es_min = config.es_min
        self.result = {"best_val": "n.a.", "num_trials": 0, "early_stopped": 0}
        if getattr(config, "seed", None) is not None:
            import random

            import numpy as np

            random.seed(config.seed)
            np.random.seed(config.seed)
        # shared state of the controller (reference optimization_driver.py:112-117)
        self.controller.num_trials = self.num_trials
        self.controller.searchspace = self.searchspace
        self.controller.trial_store = self._trial_store
        self.controller.final_store = self._final_store
        self.controller.direction = self.direction
        self.controller._initialize(exp_dir=self.log_dir)
        # scheduling state
        self._assigned = {}          # worker -> trial_id currently assigned / running
        self._idle = {}              # worker -> time of the last failed suggestion
        self._scheduled = 0
        self._controller_exhausted = False
        self._lookahead_enabled = isinstance(self.controller, (GP, TPE)) and not self.controller.pruner
        self._lookahead = None       # a Trial computed ahead of time
        self._lookahead_thread: Optional[threading.Thread] = None
        self._controller_lock = threading.RLock()
        self._resumable = self._load_finished_trials(getattr(config, "resume_from", None))
        self.resumed_trials = 0      # how many suggestions were answered from an earlier run's trial.json
        self.inter_trial_gaps = []   # seconds between a FINAL and the next assignment on that worker
        self.stop_latencies = []     # seconds between the decisive METRIC record and the STOP word

    # ------------------------------------------------------------------ hooks
    def _exp_startup_callback(self) -> None:
        tracing.start(os.path.join(self.log_dir, "trace_driver.json"), "driver")
        try:
            tensorboard._register(self.log_dir)
            tensorboard._write_hparams_config(self.log_dir, self.searchspace)
        except Exception:
            pass

    def _exp_final_callback(self, job_end: float, exp_json: dict) -> dict:
        result = self.finalize(job_end)
        best_logdir = os.path.join(self.log_dir, str(result.get("best_id")))
        util.finalize_experiment(exp_json, float(result["best_val"]) if result.get("best_id") else None, self.app_id,
                                 self.run_id, "FINISHED", self.duration, self.log_dir, best_logdir,
                                 self.optimization_key)
        print("Finished experiment.")
        return result

    def _exp_exception_callback(self, exc: Exception) -> None:
        self.log(f"Experiment failed: {exc!r}")

    def _register_msg_callbacks(self) -> None:
        self.message_callbacks.update({
            "REG": self._register_msg_callback, "METRIC": self._metric_msg_callback,
            "FINAL": self._final_msg_callback, "LOG": self._log_msg_callback,
            "ERROR": self._error_msg_callback, "IDLE": self._idle_msg_callback, "EXIT": lambda msg: None,
        })

    def init(self, job_start: float) -> None:
        super().init(job_start)
        # only the stock median rule is mirrored natively (pump thread / K9); a subclass overriding the check is custom
        native_median = self._policy_is(self.es_policy, MedianStoppingRule) and \
            getattr(self.es_policy, "earlystop_check") is MedianStoppingRule.earlystop_check
        self._native_policy = bool(native_median)
        # MAGGY_B200_HOST_POLICY=0 leaves the device scheduler as the only native engine (tests prove K9 alone suffices)
        host_engine = self._native_policy and os.environ.get("MAGGY_B200_HOST_POLICY", "1") != "0"
        self.channel.configure_policy(host_engine, self.direction, self.es_min, self.es_interval)
        self.device_scheduler = None
        self.device_scheduler_stats = None
        # ASHA rung bookkeeping + the promote / fresh-configuration decision on the device (K9); the driver samples
        self._device_asha = False
        self._dec_expected = {}
        self._staged = {}              # worker -> prefetched Trial parked in its staged slot
        self.device_handoffs = 0
        self.device_decisions = 0
        self.device_decision_fallbacks = 0
        self.device_decision_latencies = []
        self._start_device_scheduler()

    def _start_device_scheduler(self) -> None:
        """K9: when the workers run on GPUs and the policy is device-decidable, a persistent scheduler kernel takes
        the early-stop decisions (reference mechanism: heartbeat METRIC -> STOP reply, maggy/core/rpc.py:716-737,
        447-461) out of a peer-visible mailbox in its own GPU's memory; the host engines keep running — STOP is
        idempotent.

        Placement: a persistent kernel in the DRIVER's CUDA context time-slices with whatever context trains on the
        same GPU (no MPS here: contexts of different processes never run side by side), which would halve that
        worker's throughput. ``MAGGY_B200_DEVICE_SCHED=auto`` (default) therefore starts K9 only when a GPU is left
        over after the workers took theirs (GPU ``num_executors * gpus_per_worker``); ``=1`` forces it onto cuda:0
        (used by the tests and by single-GPU debugging), ``=0`` disables it."""
        mode = os.environ.get("MAGGY_B200_DEVICE_SCHED", "auto").lower()
        from ...optimizer.asha import Asha

        want_asha = isinstance(self.controller, Asha) and type(self.controller).get_suggestion is Asha.get_suggestion \
            and os.environ.get("MAGGY_B200_DEVICE_ASHA", "1") != "0" and not self._resumable
        if not (self._native_policy or want_asha) or mode in ("0", "off", "false") or self.num_executors > 32:
            return
        try:
            import torch

            ngpu = util.num_physical_devices() if torch.cuda.is_available() else 0
            taken = self.num_executors * max(1, getattr(self, "gpus_per_worker", 1))
            if ngpu < 1:
                return
            if mode in ("1", "on", "true", "force"):
                dev = min(taken, ngpu - 1) if ngpu > taken else 0
            elif ngpu > taken:
                dev = taken
            else:
                self.log("device scheduler (K9) not started: no GPU left over after the workers' "
                         f"({ngpu} visible, {taken} taken); the native host engine decides. MAGGY_B200_DEVICE_SCHED=1 forces it.")
                return
            from ..sched_kernel import DeviceScheduler

            asha = None
            if want_asha and self.controller.max_rung < 8:
                asha = {"rf": self.controller.reduction_factor, "max_rung": self.controller.max_rung,
                        "cap": max(64, 2 * int(self.num_trials))}
            # without a device-decidable early-stop policy K9 only schedules: an unreachable es_min disarms the rule
            es_min = self.es_min if self._native_policy else 0x7fffffff
            self.device_scheduler = DeviceScheduler(self.channel, direction=self.direction, es_min=es_min,
                                                    es_interval=self.es_interval, device=f"cuda:{dev}",
                                                    asha=asha).start()
            self._device_asha = asha is not None
            self.log(f"device scheduler (K9) running on cuda:{dev}, "
                     f"mailbox={'yes' if self.device_scheduler.mailbox else 'no'}, "
                     f"asha_on_device={'yes' if self._device_asha else 'no'}")
        except Exception as e:  # noqa: BLE001 - the host policy engine covers every decision on its own
            self.device_scheduler = None
            self._device_asha = False
            self.log(f"device scheduler not started: {type(e).__name__}: {e}")

    def stop(self) -> None:
        self._stop_device_scheduler()  # the kernel polls the segment: it must retire before the segment is unmapped
        super().stop()

    def _stop_device_scheduler(self) -> None:
        if getattr(self, "device_scheduler", None) is not None:
            try:
                self.device_scheduler_stats = self.device_scheduler.stop()
                self.log(f"device scheduler (K9) stats: {self.device_scheduler_stats}")
            except Exception as e:  # noqa: BLE001
                self.log(f"device scheduler stop failed: {e}")
            self.device_scheduler = None

    def _picklable_config(self, config):
        # the controller / searchspace objects stay on the driver; workers only need these fields
        class _Slim:
            pass

        slim = _Slim()
        for k in ("name", "description", "hb_interval", "model", "dataset", "optimization_key"):
            setattr(slim, k, getattr(config, k, None))
        slim.optimization_key = self.optimization_key
        return slim

    # ------------------------------------------------------------------ controller access
    def controller_get_next(self, trial=None):
        return self.controller.get_suggestion(trial)

    def _device_suggest(self, worker, finished_trial):
        """K9 filed ``finished_trial`` in its rung and decided what ``worker`` runs next; apply that decision. Returns
        ``(handled, suggestion)``; not handled = no decision arrived in time and the host controller decides."""
        expected = self._dec_expected.get(worker, 0)
        dec = self.channel.device_decision(worker, expected) if expected > 0 else None
        if dec is None and expected > 0 and worker in self._staged:
            # A trial is parked for this worker, so K9 may still publish it: the driver must not write the assign words
            # at the same time. K9 answers within microseconds unless its context is time-sliced against a trainer on
            # the same GPU (forced single-GPU mode); give it 250 ms, then retire the device path for good.
            dec = self.channel.device_decision(worker, expected, timeout_s=0.25)
            if dec is None:
                self.log("device scheduler did not answer within 250 ms: next-trial decisions return to the driver")
                self._device_asha = False
        if dec is None:
            self.device_decision_fallbacks += 1
            return False, None
        kind, rung, parent, t_ns = dec
        self.device_decisions += 1
        if getattr(finished_trial, "_final_t_ns", None):
            self.device_decision_latencies.append((t_ns - finished_trial._final_t_ns) / 1e6)
        if kind == 3:   # MG_DEC_NEW_PUBLISHED: K9 already handed the worker its prefetched trial
            state, tid = self.channel.staged_state(worker)
            parked = self._staged.pop(worker, None)
            if state == "published" and parked is not None and parked.trial_id == tid:
                parked._handoff_t_ns = t_ns
                return True, parked
            kind = 1    # (cannot happen unless the slot was re-staged in between): sample a fresh one the normal way
        with self._controller_lock:
            return True, self.controller.apply_device_decision(kind, rung, parent, finished_trial)

    def _stage_prefetch(self, worker: int) -> None:
        """Park a fresh rung-0 configuration in ``worker``'s staged slot: K9 publishes it (five words and a release
        store from the GPU) the moment its rung scan says "nothing to promote" for that worker's next FINAL."""
        if not self._device_asha or worker in self._staged or self.experiment_done or self._controller_exhausted:
            return
        if self._scheduled + len(self._staged) >= self.num_trials:
            return
        with self._controller_lock:
            if self.controller.max_rung in self.controller.rungs:
                return
            trial = self.controller.fresh_trial()
        payload = {"params": trial.params, "trial_type": trial.trial_type}
        for _ in range(40):
            # refused while the worker is still copying a hand-off out of the staged area (microseconds)
            if self.channel.stage(worker, trial.trial_id, payload, self.controller.rung_of(trial) + 1):
                trial._device_tagged = True
                self._staged[worker] = trial
                return
            time.sleep(0.0001)
        self._parked_orphans = getattr(self, "_parked_orphans", 0) + 1   # sampled but never parked: rung 0 ignores it

    def _suggest(self, finished_trial=None, worker=None):
        """Next Trial | "IDLE" | None. Uses the lookahead buffer when it holds something."""
        if self._device_asha and finished_trial is not None and worker is not None \
                and getattr(finished_trial, "_device_tagged", False):
            handled, nxt = self._device_suggest(worker, finished_trial)
            if handled:
                return nxt
        if self._lookahead_enabled:
            t = self._lookahead_thread
            if t is not None and not t.is_alive():
                self._lookahead_thread = None
            if self._lookahead is not None:
                nxt, self._lookahead = self._lookahead, None
                return nxt
            if self._lookahead_thread is not None:
                self._lookahead_thread.join()
                self._lookahead_thread = None
                if self._lookahead is not None:
                    nxt, self._lookahead = self._lookahead, None
                    return nxt
        with self._controller_lock:
            return self.controller_get_next(finished_trial)

    def _kick_lookahead(self, finished_trial):
        if not self._lookahead_enabled or self._lookahead is not None or self._lookahead_thread is not None:
            return
        if self._controller_exhausted or self._scheduled + 1 >= self.num_trials + len(self._idle):
            return

        def work():
            try:
                with self._controller_lock:
                    nxt = self.controller_get_next(finished_trial)
                if isinstance(nxt, Trial):
                    self._lookahead = nxt
            except Exception as exc:  # noqa: BLE001
                self.exception = exc

        self._lookahead_thread = threading.Thread(target=work, name="maggy-suggest", daemon=True)
        self._lookahead_thread.start()

    # ------------------------------------------------------------------ scheduling
    def _assign_next(self, worker: int, finished_trial=None, t_final=None):
        if self.experiment_done:
            return
        if self._controller_exhausted:
            self._maybe_finish()
            return
        if self._lookahead_enabled and self._scheduled >= self.num_trials:
            # model-based controllers count *finished* trials; with several workers that would overshoot
            # num_trials by the number of trials in flight
            self._controller_exhausted = True
            self._maybe_finish()
            return
        nxt = self._suggest(finished_trial, worker)
        if nxt is None:
            self._controller_exhausted = True
            self._idle.pop(worker, None)
            self._maybe_finish()
            return
        if nxt == "IDLE":
            self._idle[worker] = (time.time(), finished_trial)
            return
        self._idle.pop(worker, None)
        if nxt.trial_id in self._resumable:
            # experiment-level resume: this configuration already finished in the run we are resuming from
            done = self._replay_finished(nxt)
            self._assign_next(worker, done, t_final)
            return
        handoff_t = getattr(nxt, "_handoff_t_ns", None)
        if handoff_t is not None:
            # the device scheduler published this (prefetched) trial already: bookkeeping only
            self._adopt_published(worker, nxt)
            if getattr(finished_trial, "_final_t_ns", None):
                self.inter_trial_gaps.append(max(0.0, (handoff_t - finished_trial._final_t_ns) / 1e9))
        else:
            self._schedule(worker, nxt)
            if t_final is not None:
                self.inter_trial_gaps.append(time.time() - t_final)
        self._stage_prefetch(worker)
        self._kick_lookahead(finished_trial)

    def _adopt_published(self, worker: int, trial: Trial):
        with self._store_lock:
            self._trial_store[trial.trial_id] = trial
        trial.start = time.time()
        trial.status = Trial.RUNNING
        self._assigned[worker] = trial.trial_id
        self._scheduled += 1
        self.device_handoffs += 1
        self.channel.reservations.assign_trial(worker, trial.trial_id)
        tracing.instant("assign", worker=worker, trial=trial.trial_id, handoff="device")

    def _schedule(self, worker: int, trial: Trial):
        with self._store_lock:
            if trial.trial_id in self._trial_store:
                self.log(f"WARNING trial {trial.trial_id} is already running; asking the controller again")
            self._trial_store[trial.trial_id] = trial
        trial.status = Trial.SCHEDULED
        trial.start = time.time()
        payload = {"params": trial.params, "trial_type": trial.trial_type}
        tag = 0
        if self._device_asha:
            tag = self.controller.rung_of(trial) + 1   # echoed in the FINAL record: K9 files the result in this rung
            trial._device_tagged = True
        while not self.channel.assign(worker, trial.trial_id, payload, tag):
            time.sleep(0.0005)  # the worker has not picked up its previous assignment yet
        parked = self._staged.pop(worker, None)
        if parked is not None:
            # a direct assignment invalidates the parked entry (its sequence number is taken): park it again
            if self.channel.stage(worker, parked.trial_id, {"params": parked.params, "trial_type": parked.trial_type},
                                  self.controller.rung_of(parked) + 1):
                self._staged[worker] = parked
        trial.status = Trial.RUNNING
        self._assigned[worker] = trial.trial_id
        self._scheduled += 1
        tracing.instant("assign", worker=worker, trial=trial.trial_id)

    # ------------------------------------------------------------------ resume (additive)
    def _load_finished_trials(self, path) -> dict:
        """``{trial_id: Trial}`` for every FINALIZED ``<path>/<trial_id>/trial.json`` of an earlier run."""
        out = {}
        if not path:
            return out
        import glob

        for f in glob.glob(os.path.join(path, "*", "trial.json")):
            try:
                t = Trial.from_json(open(f).read())
            except Exception:  # noqa: BLE001 - a torn file from a crashed run is simply not resumable
                continue
            if t.status == Trial.FINALIZED and t.final_metric is not None:
                out[t.trial_id] = t
        self.log(f"resume_from={path}: {len(out)} finished trials can be replayed")
        return out

    def _replay_finished(self, suggested: Trial) -> Trial:
        saved = self._resumable.pop(suggested.trial_id)
        with self._controller_lock, self._store_lock:
            suggested.status = Trial.FINALIZED
            suggested.final_metric = saved.final_metric
            suggested.metric_history = list(saved.metric_history)
            suggested.step_history = list(saved.step_history)
            suggested.metric_dict = dict(saved.metric_dict)
            suggested.early_stop = saved.early_stop
            suggested.duration = saved.duration
            self._final_store.append(suggested)
        self._scheduled += 1
        self.resumed_trials += 1
        self._update_result(suggested)
        self.maggy_log = self._update_maggy_log()
        self.log(f"replayed trial {suggested.trial_id} from the resumed run: {self.maggy_log}")
        EnvSing.get_instance().dump(suggested.to_json(), os.path.join(self.log_dir, suggested.trial_id, "trial.json"))
        return suggested

    def _maybe_finish(self):
        if not self._assigned and (self._controller_exhausted or len(self._final_store) >= self.num_trials):
            self.experiment_done = True

    def _on_tick(self) -> None:
        if not self._idle or self.experiment_done:
            return
        now = time.time()
        for worker, (since, last) in list(self._idle.items()):
            if now - since >= 0.05:  # reference retries every 0.1 s (optimization_driver.py:550)
                self._assign_next(worker, last)

    # ------------------------------------------------------------------ message callbacks
    def _register_msg_callback(self, msg) -> None:
        worker = msg["partition_id"]
        if self._assigned.get(worker) is not None and self._blacklist_msg_callback(msg):
            return
        self._assign_next(worker, None)

    def _blacklist_msg_callback(self, msg) -> bool:
        """A restarted worker re-registered while a trial is still assigned to it: the reference marks the trial ERROR
        and re-schedules the very same trial on this partition (BLACK message, rpc.py:421-432,
        optimization_driver.py:473-483). Returns True when a lost trial was rescheduled."""
        worker = msg["partition_id"]
        lost = self._assigned.get(worker)
        trial = self.get_trial(lost) if lost is not None else None
        if trial is None:
            return False
        trial.status = Trial.ERROR
        self.log(f"worker {worker} re-registered (attempt {msg.get('aux')}): rescheduling lost trial {lost}")
        trial.metric_history, trial.step_history, trial.metric_dict = [], [], {}
        trial.early_stop = False
        self._scheduled -= 1
        self._schedule(worker, trial)
        return True

    def _idle_msg_callback(self, msg) -> None:
        """A worker without work asks again (reference: optimization_driver.py:540-553 retries every 0.1 s)."""
        worker = msg["partition_id"]
        if worker in self._idle and not self.experiment_done:
            self._assign_next(worker, self._idle[worker][1])

    def _metric_msg_callback(self, msg) -> None:
        trial = self.get_trial(msg["trial_id"])
        if trial is None:
            return
        new_step = trial.append_metric({"value": msg["value"], "step": msg["step"]})
        if msg["stopped"]:
            trial.set_early_stop()
            self.stop_latencies.append(max(0.0, (self.channel.stop_time_ns(msg["partition_id"]) - msg["t_ns"]) / 1e9))
            return
        if self._native_policy or self._no_earlystop or new_step is None or trial.get_early_stop():
            return
        with self._store_lock:
            finals = list(self._final_store)
        if len(finals) > self.es_min and new_step % self.es_interval == 0:
            try:
                to_stop = self.earlystop_check(trial, finals, self.direction)
            except Exception as e:  # noqa: BLE001
                self.log(str(e))
                to_stop = None
            if to_stop is not None:
                self.log(f"Trials to stop: {to_stop}")
                trial.set_early_stop()
                self.channel.stop_trial(msg["partition_id"], trial.trial_id)

    def _log_msg_callback(self, msg) -> None:
        if msg["data"]:
            self.add_executor_logs(msg["data"])

    def _error_msg_callback(self, msg) -> None:
        text = (msg["data"] or {}).get("error", "unknown error") if isinstance(msg["data"], dict) else str(msg["data"])
        trial = self.get_trial(msg["trial_id"]) if msg["trial_id"] else None
        if trial is not None:
            trial.status = Trial.ERROR
        self.log(f"worker {msg['partition_id']} failed in trial {msg['trial_id']}:\n{text}")
        self.exception = RuntimeError(
            f"train_fn raised in trial {msg['trial_id']} on worker {msg['partition_id']}:\n{text}")

    def _final_msg_callback(self, msg) -> None:
        t_final = time.time()
        worker = msg["partition_id"]
        trial = self.get_trial(msg["trial_id"])
        data = msg["data"] or {}
        if trial is None:
            return
        if data.get("logs"):
            self.add_executor_logs(data["logs"])
        if data.get("early_stop") and not trial.get_early_stop():
            # the STOP word was written by an engine the digest thread did not hear from (the device scheduler K9)
            trial.set_early_stop()
        with self._controller_lock, self._store_lock:
            trial.status = Trial.FINALIZED
            trial.final_metric = data.get("metric", msg["value"])
            trial.duration = util.seconds_to_milliseconds(time.time() - trial.start)
            self._final_store.append(trial)
            self._trial_store.pop(trial.trial_id, None)
        self._assigned.pop(worker, None)
        if getattr(trial, "_device_tagged", False):
            # K9 takes exactly one next-trial decision per tagged FINAL of this worker
            self._dec_expected[worker] = self._dec_expected.get(worker, 0) + 1
            trial._final_t_ns = msg.get("t_ns")
        tracing.instant("final", worker=worker, trial=trial.trial_id, metric=trial.final_metric)
        # hand the worker its next trial first, do the bookkeeping while it is already training
        self._assign_next(worker, trial, t_final)
        self._update_result(trial)
        self.maggy_log = self._update_maggy_log()
        self.log(self.maggy_log)
        EnvSing.get_instance().dump(trial.to_json(), os.path.join(self.log_dir, trial.trial_id, "trial.json"))
        self._maybe_finish()

    # ------------------------------------------------------------------ stores
    def get_trial(self, trial_id):
        with self._store_lock:
            return self._trial_store.get(trial_id)

    def add_trial(self, trial):
        with self._store_lock:
            self._trial_store[trial.trial_id] = trial

    # ------------------------------------------------------------------ results / artefacts
    def _update_result(self, trial: Trial) -> None:
        metric = trial.final_metric
        params = {k: v for k, v in trial.params.items() if k not in ("dataset_function", "model_function")}
        r = self.result
        if r.get("best_id") is None:
            self.result = {"best_id": trial.trial_id, "best_val": metric, "best_config": params,
                           "worst_id": trial.trial_id, "worst_val": metric, "worst_config": params, "avg": metric,
                           "metric_list": [metric], "num_trials": 1, "early_stopped": 1 if trial.early_stop else 0,
                           "num_epochs": len(trial.metric_history), "trial_id": trial.trial_id}
            return
        better = (lambda a, b: a > b) if self.direction == "max" else (lambda a, b: a < b)
        if better(metric, r["best_val"]):
            r.update(best_val=metric, best_id=trial.trial_id, best_config=params)
        if better(r["worst_val"], metric):
            r.update(worst_val=metric, worst_id=trial.trial_id, worst_config=params)
        r["metric_list"].append(metric)
        r["num_trials"] += 1
        r["avg"] = sum(r["metric_list"]) / float(len(r["metric_list"]))
        if trial.early_stop:
            r["early_stopped"] += 1

    def finalize(self, job_end: float) -> dict:
        self._stop_device_scheduler()
        self.job_end = job_end
        self.duration = util.seconds_to_milliseconds(job_end - self.job_start)
        results = self.prep_results(util.time_diff(self.job_start, job_end))
        print(results)
        self.log(results)
        env = EnvSing.get_instance()
        env.dump(json.dumps(self.result, default=util.json_default_numpy), os.path.join(self.log_dir, "result.json"))
        env.dump(self.json(), os.path.join(self.log_dir, "maggy.json"))
        return self.result

    def prep_results(self, duration_str: str) -> str:
        self.controller._finalize_experiment(self._final_store)
        r = self.result
        return ("\n------ " + self.controller.name() + " Results ------ direction(" + self.direction + ") \n"
                "BEST combination " + json.dumps(r.get("best_config"), default=util.json_default_numpy)
                + " -- metric " + str(r.get("best_val")) + "\n"
                "WORST combination " + json.dumps(r.get("worst_config"), default=util.json_default_numpy)
                + " -- metric " + str(r.get("worst_val")) + "\n"
                "AVERAGE metric -- " + str(r.get("avg")) + "\n"
                "EARLY STOPPED Trials -- " + str(r.get("early_stopped")) + "\n"
                "Total job time " + duration_str + "\n")

    def config_to_dict(self) -> dict:
        return self.searchspace.to_dict()

    def json(self) -> str:
        env = EnvSing.get_instance()
        exp = {
            "project": env.project_name(), "user": env.get_user() or None, "name": self.name, "module": "maggy",
            "app_id": str(self.app_id),
            "start": time.strftime("%Y-%m-%dT%H:%M:%S", time.localtime(self.job_start)),
            "memory_per_executor": str(os.environ.get("MAGGY_EXECUTOR_MEMORY")),
            "gpus_per_executor": str(self.gpus_per_worker if util.num_physical_devices() else 0),
            "executors": self.num_executors, "logdir": self.log_dir, "description": self.description,
            "experiment_type": self.controller.name(), "controller": self.controller.name(),
            "config": json.dumps(self.config_to_dict(), default=util.json_default_numpy),
        }
        if self.experiment_done:
            exp["status"] = "FINISHED"
            exp["finished"] = time.strftime("%Y-%m-%dT%H:%M:%S", time.localtime(self.job_end))
            exp["duration"] = self.duration
            exp["config"] = json.dumps(self.result.get("best_config"), default=util.json_default_numpy)
            exp["metric"] = self.result.get("best_val")
            if getattr(self, "device_scheduler_stats", None):
                exp["device_scheduler"] = dict(self.device_scheduler_stats)  # K9 counters: polls / metrics / stops / mailbox
                if getattr(self, "_device_asha", False) or getattr(self, "device_decisions", 0):
                    lat = sorted(getattr(self, "device_decision_latencies", []))
                    exp["device_scheduler"].update({
                        "asha_decisions_applied": self.device_decisions,
                        "prefetched_handoffs_adopted": self.device_handoffs,
                        "asha_decision_fallbacks": self.device_decision_fallbacks,
                        "final_to_decision_ms_median": lat[len(lat) // 2] if lat else None})
        else:
            exp["status"] = "RUNNING"
        return json.dumps(exp, default=util.json_default_numpy)

    def _update_maggy_log(self) -> str:
        return self.log_string()

    def log_string(self) -> str:
        r = self.result
        return ("Maggy Optimization " + str(r["num_trials"]) + "/" + str(self.num_trials) + " ("
                + str(r["early_stopped"]) + ") " + util.progress_bar(r["num_trials"], self.num_trials)
                + " - BEST " + json.dumps(r.get("best_config"), default=util.json_default_numpy)
                + " - metric " + str(r["best_val"]))

    # ------------------------------------------------------------------ validation of the config
    @staticmethod
    def _init_searchspace(searchspace):
        if searchspace is None:
            return Searchspace()
        if not isinstance(searchspace, Searchspace):
            raise Exception(f"The experiment's search space should be an instance of maggy.Searchspace, but it is "
                            f"{searchspace} (of type '{type(searchspace).__name__}').")
        return searchspace

    @staticmethod
    def _init_direction(direction):
        if not isinstance(direction, str) or direction.lower() not in ("min", "max"):
            raise Exception(f"The experiment's direction should be a string (either 'min' or 'max') but it is "
                            f"{direction} (of type '{type(direction).__name__}').")
        return direction.lower()

    def _init_controller(self, optimizer, searchspace):
        if optimizer is None:
            optimizer = "none"
        if isinstance(optimizer, str):
            key = optimizer.lower()
            if key not in self.controller_dict or self.controller_dict[key] is None:
                raise Exception(f"Unknown Optimizer. Can't initialize experiment driver. Got '{optimizer}', expected "
                                f"one of {[k for k, v in self.controller_dict.items() if v]}.")
            if key == "none" and len(searchspace.names()) > 0:
                raise Exception("Searchspace has to be empty or None to use without Optimizer.")
            return self.controller_dict[key]()
        if isinstance(optimizer, AbstractOptimizer):
            print(f"Custom Optimizer initialized: {type(optimizer).__name__}")
            return optimizer
        raise Exception(f"The experiment's optimizer should either be an string indicating the name of an implemented "
                        f"optimizer (such as 'randomsearch') or an instance of maggy.optimizer.AbstractOptimizer, but it "
                        f"is {optimizer} (of type '{type(optimizer).__name__}').")

    @staticmethod
    def _policy_is(policy, cls) -> bool:
        return isinstance(policy, cls) or (isinstance(policy, type) and issubclass(policy, cls))

    @staticmethod
    def _init_earlystop_check(es_policy):
        """Resolve ``es_policy`` to a policy class / instance whose ``earlystop_check`` the digest thread calls."""
        if isinstance(es_policy, str):
            key = es_policy.lower()
            if key == "median":
                return MedianStoppingRule
            if key == "none":
                return NoStoppingRule
            raise Exception(f"The experiment's early stopping policy should either be a string ('median' or 'none') or "
                            f"a custom policy that is an instance of maggy.earlystop.AbstractEarlyStop, but it is "
                            f"{es_policy} (of type '{type(es_policy).__name__}').")
        if isinstance(es_policy, AbstractEarlyStop) or (isinstance(es_policy, type) and issubclass(es_policy, AbstractEarlyStop)):
            print(f"Custom Early Stopping policy initialized: {getattr(es_policy, '__name__', type(es_policy).__name__)}")
            return es_policy
        raise Exception(f"The experiment's early stopping policy should either be a string ('median' or 'none') or a "
                        f"custom policy that is an instance of maggy.earlystop.AbstractEarlyStop, but it is {es_policy} "
                        f"(of type '{type(es_policy).__name__}').")
