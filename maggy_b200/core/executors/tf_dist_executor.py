"""Worker of a distributed TensorFlow training run (reference: maggy/core/executors/tf_dist_executor.py:35-480).

Flow: register -> the assignment carries this task's ``TF_CONFIG`` (cluster spec + role, built by the driver from the
registrations: R4 ``TensorflowServer``'s TF_CONFIG / RESERVATIONS messages, maggy/core/rpc.py:593-633) -> export it
-> ``MultiWorkerMirroredStrategy`` -> the model class is wrapped so that it is built under the strategy scope ->
``process_data`` + sharding of the dataset -> ``train_fn(model=, dataset=, hparams=, reporter=)`` -> FINAL.

TensorFlow is optional: the B200 compute path of this framework is PyTorch + sm_100a kernels and the image ships no
TensorFlow. Without it the worker fails the experiment with ``NotSupportedError`` and a pointer to
``TorchDistributedConfig``. ``MAGGY_B200_TF_MODULE`` names a module to import in TensorFlow's place (the test suite
uses a stand-in that provides ``distribute.MultiWorkerMirroredStrategy``).
"""
from __future__ import annotations

import builtins
import importlib
import json
import os
import time
import traceback

from ... import util
from .. import rpc
from ..exceptions import NotSupportedError
from ..reporter import Reporter
from ..tf_patching.tf_modules import get_wrapped_model
from .trial_executor import _start_heartbeat, call_train_fn


def import_tensorflow():
    name = os.environ.get("MAGGY_B200_TF_MODULE", "tensorflow")
    try:
        return importlib.import_module(name)
    except ImportError as e:
        raise NotSupportedError(
            "TfDistributedConfig", type(e).__name__,
            f"TensorFlow ('{name}') is not importable in this environment: {e}. maggy_b200's compute path is "
            "PyTorch + sm_100a kernels; use TorchDistributedConfig, or install TensorFlow to run Keras models "
            "under MultiWorkerMirroredStrategy.") from e


def shard_data(dataset, num_shards: int, index: int):
    """One worker's share (reference ``_shard_data`` :382-406): ``tf.data.Dataset.shard`` when available, every
    ``num_shards``-th element of a sequence otherwise; tuples / lists of such things are sharded element-wise."""
    if dataset is None or num_shards <= 1:
        return dataset
    if hasattr(dataset, "shard") and callable(dataset.shard):
        return dataset.shard(num_shards, index)
    if isinstance(dataset, (tuple, list)) and dataset and all(
            hasattr(d, "shard") or hasattr(d, "__getitem__") for d in dataset) and \
            not all(isinstance(d, (int, float, str)) for d in dataset):
        if all(hasattr(d, "__len__") and not isinstance(d, (str, bytes)) for d in dataset):
            return type(dataset)(shard_data(d, num_shards, index) for d in dataset)
    if hasattr(dataset, "__getitem__") and hasattr(dataset, "__len__"):
        return dataset[index::num_shards]
    return dataset


def consume_data(config):
    """``process_data(dataset)`` if the user supplied a pre-processing function (reference ``_consume_data``)."""
    data = config.dataset
    if data is not None and callable(getattr(config, "process_data", None)):
        data = config.process_data(data)
    return data


def tf_dist_executor_loop(spec, channel, worker_id, attempt):
    config = spec["config"]
    log_dir = os.path.join(spec["log_dir"], f"training_logs_{worker_id}")
    os.makedirs(log_dir, exist_ok=True)
    exec_log = os.path.join(spec["log_dir"], f"executor_{worker_id}.log")
    original_print = builtins.print
    reporter = Reporter(exec_log, worker_id, attempt, original_print, channel)
    reporter.init_logger(os.path.join(log_dir, "output.log"))

    def maggy_print(*args, **kwargs):
        original_print(*args, **kwargs)
        reporter.log(" ".join(str(x) for x in args), True)

    builtins.print = maggy_print
    hb_stop = _start_heartbeat(channel, max(0.2, float(spec.get("hb_interval", 1))))
    clean_exit = False
    try:
        channel.register()
        got = None
        while got is None:
            got = channel.wait_assign(timeout_s=1.0)
        if got is rpc.WorkerChannel.GSTOP:
            clean_exit = True
            return
        trial_id, payload = got
        channel.set_state(rpc.W_RUNNING)
        reporter.set_trial_id(trial_id)
        tf_config = payload["tf_config"]
        os.environ["TF_CONFIG"] = json.dumps(tf_config)
        reporter.log(f"Tensorflow config is {tf_config}", False)
        tf = import_tensorflow()
        if getattr(config, "mixed_precision", False):
            try:  # bf16 is the native mixed-precision type of the B200 tensor cores
                tf.keras.mixed_precision.set_global_policy("mixed_bfloat16")
            except AttributeError:
                pass
        role = tf_config["task"]["type"]
        strategy = tf.distribute.MultiWorkerMirroredStrategy()
        model = get_wrapped_model(config.model, strategy, role == "chief")
        data = shard_data(consume_data(config), payload["num_training_tasks"], payload["training_index"]) \
            if role != "evaluator" else consume_data(config)
        t0 = time.time()
        retval = call_train_fn(spec["train_fn"], {"model": model, "dataset": data, "hparams": config.hparams,
                                                  "reporter": reporter})
        if isinstance(retval, dict):       # several metrics: the first one is the experiment's "Metric"
            retval = next(iter(retval.values())) if retval else None
        elif isinstance(retval, (list, tuple)):
            retval = retval[0] if retval else None
        metric = util.handle_return_val({"Metric": retval}, log_dir, "Metric", os.path.join(log_dir, "output.log"))
        _, _, logs = reporter.get_data()
        channel.post_final(trial_id, metric, {"logs": logs, "wall_s": time.time() - t0, "role": role})
        clean_exit = True
    except Exception:
        text = traceback.format_exc()
        reporter.log(text, False)
        channel.post_error(None, text)
        raise
    finally:
        hb_stop.set()
        builtins.print = original_print
        try:
            channel.set_state(rpc.W_DONE if clean_exit else rpc.W_DEAD)
            if clean_exit:
                channel.post_exit()
        except Exception:
            pass
        reporter.close_logger()



def dist_executor_fn(train_fn, config, experiment_type, app_id, run_id, server_addr, hb_interval, secret, optimization_key, log_dir):
    """The reference's executor factory (same argument list): ``executor_fn(partition_id)`` that attaches to the control
    segment ``server_addr`` and runs this module's loop — see ``worker_main.make_executor_fn``."""
    from .worker_main import make_executor_fn

    return make_executor_fn("tf_dist", train_fn, config, experiment_type, app_id, run_id, server_addr, hb_interval, secret,
                            optimization_key, log_dir)



def find_index(host_port, reservations):
    """Index of ``host_port`` among the cluster's workers, -1 if absent (reference: tf_dist_executor.py:467-480)."""
    workers = reservations["cluster"]["worker"]
    return workers.index(host_port) if host_port in workers else -1
