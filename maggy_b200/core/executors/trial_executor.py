"""Worker loop for hyperparameter-optimization and ablation trials
(reference: maggy/core/executors/trial_executor.py:35-213).

Per trial: create/clean ``<log_dir>/<trial_id>/``, open the trial log, register the tensorboard dir,
dump ``.hparams.json``, call ``train_fn`` with the arguments its signature asks for
(``model, dataset, hparams, reporter``), validate the return value (``.outputs.json``, ``.metric``),
send FINAL. Between trials the worker futex-waits on its assign slot instead of polling ``GET`` once
per second (reference rpc.py:739-748).
"""
from __future__ import annotations

import builtins
import inspect
import json
import os
import threading
import time
import traceback

from ... import tensorboard, util
from .. import exceptions, rpc
from ..environment.singleton import EnvSing
from ..reporter import Reporter

_STOP = "__maggy_stop__"


def _start_heartbeat(channel, interval_s: float, reporter=None):
    stop = threading.Event()

    def beat():
        while not stop.wait(interval_s):
            try:
                channel.heartbeat()
                if reporter is not None and reporter.deferred:
                    reporter.ship()
            except Exception:
                return

    t = threading.Thread(target=beat, name="maggy-heartbeat", daemon=True)
    t.start()
    return stop


def _preimport():
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        pass


from ...utils import tracing  # noqa: E402

def call_train_fn(train_fn, available: dict):
    """Inject exactly the keyword arguments train_fn declares (reference trial_executor.py:166-179)."""
    sig = inspect.signature(train_fn)
    if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values()):
        return train_fn(**available)
    return train_fn(**{k: v for k, v in available.items() if k in sig.parameters})


def run_one_trial(spec, trial_id, params, reporter, log_dir, gang=None):
    """Execute train_fn for one trial; returns the numeric optimization metric."""
    env = EnvSing.get_instance()
    train_fn = spec["train_fn"]
    config = spec["config"]
    trial_dir = os.path.join(log_dir, trial_id)
    trial_log = os.path.join(trial_dir, "output.log")
    leader = gang is None or gang.is_leader
    if leader:
        if env.exists(trial_dir):
            util.clean_dir(trial_dir, keep=["output.log"])  # a rescheduled trial reuses its directory
        else:
            env.mkdir(trial_dir)
        reporter.init_logger(trial_log)
    tensorboard._register(trial_dir)
    if gang is not None:
        reporter.gang = gang
        gang.reset_agreement()
    hparams = dict(params)
    if spec["experiment_type"] == "ablation":
        printable = {k: hparams.pop(k) for k in ("ablated_feature", "ablated_layer") if k in hparams}
    else:
        printable = hparams
    if leader:
        env.dump(json.dumps(printable, default=util.json_default_numpy), os.path.join(trial_dir, ".hparams.json"))
        if spec["experiment_type"] != "ablation":
            try:
                tensorboard._write_hparams(printable, trial_id)
            except Exception:
                pass
    reporter.log(f"Starting Trial: {trial_id}", False)
    reporter.log(f"Trial Configuration: {printable}", False)
    available = {"model": getattr(config, "model", None), "dataset": getattr(config, "dataset", None),
                 "hparams": hparams, "reporter": reporter}
    try:
        with tracing.range(f"trial {trial_id}", params={k: repr(v)[:60] for k, v in printable.items()}):
            retval = call_train_fn(train_fn, available)
        if isinstance(retval, dict) and spec["optimization_key"] not in retval and len(retval) >= 1 \
                and spec["optimization_key"] == "Metric":
            # reference: with the default key the first value of the returned dict is the metric
            retval = {"Metric": next(iter(retval.values())), **retval}
        if leader:
            retval = util.handle_return_val(retval, trial_dir, spec["optimization_key"], trial_log)
        elif isinstance(retval, dict):
            retval = retval[spec["optimization_key"]]
    except exceptions.EarlyStopException as e:
        retval = e.metric
        reporter.early_stopped = True   # travels with FINAL: whoever wrote the STOP word (host engine, K9, driver)
        reporter.log("Early Stopped Trial.", False)
    if gang is not None:
        retval = gang.mean(retval)
    reporter.log(f"Finished Trial: {trial_id}", False)
    reporter.log(f"Final Metric: {retval}", False)
    return retval


def trial_executor_loop(spec, channel, worker_id, attempt, gang=None):
    tracing.start(os.path.join(spec["log_dir"], f"trace_worker{worker_id}.json"), f"worker{worker_id}")
    env = EnvSing.get_instance()
    log_dir = spec["log_dir"]
    if gang is not None and not gang.is_leader:
        return _follower_loop(spec, worker_id, attempt, gang, log_dir)

    exec_log = os.path.join(log_dir, f"executor_{worker_id}_{attempt}.log")
    original_print = builtins.print
    reporter = Reporter(exec_log, worker_id, attempt, original_print, channel)

    def maggy_print(*args, **kwargs):
        original_print(*args, **kwargs)
        reporter.log(" ".join(str(x) for x in args), True)

    builtins.print = maggy_print
    proxy_rpc = os.environ.get("MAGGY_B200_PROXY_REFERENCE_RPC") == "1"
    reporter.deferred = proxy_rpc
    hb_stop = _start_heartbeat(channel, float(spec.get("hb_interval", 1)) if proxy_rpc
                               else max(0.2, float(spec.get("hb_interval", 1))), reporter)
    clean_exit = False
    try:
        channel.register()
        reporter.log(f"Registered with experiment driver (segment {channel.name})", False)
        _preimport()  # pay the torch / tensorboard import once, while the driver prepares the first trial
        while True:
            channel.set_state(rpc.W_IDLE)
            got = channel.wait_assign(timeout_s=1.0)
            if got is rpc.WorkerChannel.GSTOP:
                clean_exit = True
                break
            if got is None:
                continue
            trial_id, payload = got
            channel.set_state(rpc.W_RUNNING)
            reporter.set_trial_id(trial_id)
            if gang is not None:
                gang.broadcast((trial_id, payload))
            t0 = time.time()
            try:
                metric = run_one_trial(spec, trial_id, payload["params"], reporter, log_dir, gang)
            except Exception:
                text = traceback.format_exc()
                reporter.log(text, False)
                channel.post_error(trial_id, text)
                reporter.reset()
                raise
            _, _, logs = reporter.get_data()
            channel.post_final(trial_id, metric, {"logs": logs, "wall_s": time.time() - t0,
                                                  "early_stop": bool(getattr(reporter, "early_stopped", False))})
            reporter.early_stopped = False
            reporter.reset()
            if proxy_rpc:
                time.sleep(1.0)  # reference: the first GET after FINAL races the driver and sleeps 1 s (rpc.py:745-747)
    finally:
        hb_stop.set()
        builtins.print = original_print
        if gang is not None:
            try:
                gang.broadcast(_STOP)
            except Exception:
                pass
        try:
            # W_DONE only after an orderly global stop: anything else is a death the driver must handle
            channel.set_state(rpc.W_DONE if clean_exit else rpc.W_DEAD)
            if clean_exit:
                channel.post_exit()
        except Exception:
            pass
        reporter.close_logger()


def _follower_loop(spec, worker_id, attempt, gang, log_dir):
    """Gang ranks > 0: run the same trials as the leader; only the leader talks to the driver."""
    reporter = Reporter(None, worker_id, attempt, None, None)
    while True:
        msg = gang.broadcast(None)
        if msg == _STOP or msg is None:
            return
        trial_id, payload = msg
        reporter.set_trial_id(trial_id)
        run_one_trial(spec, trial_id, payload["params"], reporter, log_dir, gang)
        reporter.reset()



def trial_executor_fn(train_fn, config, experiment_type, app_id, run_id, server_addr, hb_interval, secret, optimization_key, log_dir):
    """The reference's executor factory (same argument list): ``executor_fn(partition_id)`` that attaches to the control
    segment ``server_addr`` and runs this module's loop — see ``worker_main.make_executor_fn``."""
    from .worker_main import make_executor_fn

    return make_executor_fn("trial", train_fn, config, experiment_type, app_id, run_id, server_addr, hb_interval, secret,
                            optimization_key, log_dir)
