"""Helpers shared by drivers and executors (reference: maggy/util.py:39-365)."""
from __future__ import annotations

import json
import math
import os
import time
from typing import Optional

import numpy as np

from . import constants
from .core import exceptions

DEBUG = bool(int(os.environ.get("MAGGY_DEBUG", "0")))
APP_ID = None
RUN_ID = 1


def log(msg):
    if DEBUG:
        print(msg)


def json_default_numpy(obj):
    """``json.dumps(default=...)`` hook: numpy scalars/arrays and torch scalars become plain numbers."""
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if hasattr(obj, "item") and callable(obj.item) and getattr(obj, "ndim", 1) == 0:
        return obj.item()
    if callable(obj):
        return getattr(obj, "__name__", repr(obj))
    raise TypeError(f"Object of type {type(obj)}: {obj} is not JSON serializable")


def progress_bar(done: int, total: int) -> str:
    """``[=====>........]`` — 30 cells, the cell at the progress front is '>'."""
    front = math.floor(done / total * 30) if total else 30
    return "[" + "".join("=" if i < front else (">" if i == front else ".") for i in range(30)) + "]"


def time_diff(t0: float, t1: float) -> str:
    """Human readable ``t1 - t0`` (seconds since epoch), in the reference's fixed format (maggy/util.py:263-272)."""
    minutes, seconds = divmod(max(0.0, t1 - t0), 60)
    hours, minutes = divmod(minutes, 60)
    return "%d hours, %d minutes, %d seconds" % (hours, minutes, seconds)


def seconds_to_milliseconds(t: float) -> int:
    return int(round(t * 1000))


def get_metric_value(return_dict, metric_key):
    """The optimization metric out of a train_fn return dict (reference: maggy/util.py:350-365): ``return_dict[key]``
    when a key is given, the ``"metric"`` entry of a two-entry dict otherwise, else None."""
    if return_dict and metric_key:
        assert metric_key in return_dict.keys(), f"Supplied metric_key {metric_key} is not in returned dict {return_dict}"
        return return_dict[metric_key]
    if return_dict is not None and len(return_dict.keys()) == 2 and "metric" in return_dict.keys():
        return return_dict["metric"]
    return None


def num_physical_devices() -> int:
    """Accelerators visible to this process (0 on a CPU-only machine)."""
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def num_executors(default: Optional[int] = None) -> int:
    """Size of the worker pool: MAGGY_NUM_EXECUTORS, else one worker per visible GPU, else ``default``
    (CPU plumbing mode), else 1."""
    env = os.environ.get("MAGGY_NUM_EXECUTORS")
    if env:
        return max(1, int(env))
    n = num_physical_devices()
    if n:
        return n
    return max(1, default or 1)


def clean_dir(path: str, keep=()):
    """Delete everything in ``path`` except entries named in ``keep`` (trial dir reuse on retry)."""
    if not os.path.isdir(path):
        return
    import shutil

    for entry in os.listdir(path):
        if entry in keep:
            continue
        full = os.path.join(path, entry)
        if os.path.isdir(full):
            shutil.rmtree(full, ignore_errors=True)
        else:
            try:
                os.remove(full)
            except OSError:
                pass


def handle_return_val(return_val, log_dir: str, optimization_key: str, log_file: str):
    """Validate what train_fn returned and persist ``.outputs.json`` + ``.metric`` in the trial dir.

    Accepted: a number, or a dict containing ``optimization_key`` with a numeric value
    (reference: maggy/util.py:159-199). Returns the numeric optimization value.
    """
    from .core.environment.singleton import EnvSing

    env = EnvSing.get_instance()
    if not optimization_key:
        raise ValueError("Optimization key cannot be None.")
    if return_val is None or (isinstance(return_val, dict) and not return_val):
        raise exceptions.ReturnTypeError(optimization_key, return_val)
    if hasattr(return_val, "item") and getattr(return_val, "ndim", 1) == 0:
        return_val = return_val.item()  # 0-dim tensor / numpy scalar
    if isinstance(return_val, bool) or not isinstance(return_val, constants.USER_FCT.RETURN_TYPES):
        raise exceptions.ReturnTypeError(optimization_key, return_val)
    if isinstance(return_val, dict):
        if optimization_key not in return_val:
            raise KeyError(
                f"Returned dictionary does not contain optimization key with the provided name: {optimization_key}")
        outputs = dict(return_val)
        opt_val = outputs[optimization_key]
        if hasattr(opt_val, "item") and getattr(opt_val, "ndim", 1) == 0:
            opt_val = outputs[optimization_key] = opt_val.item()
    else:
        opt_val = return_val
        outputs = {optimization_key: opt_val}
    if isinstance(opt_val, bool) or not isinstance(opt_val, constants.USER_FCT.NUMERIC_TYPES):
        raise exceptions.MetricTypeError(optimization_key, opt_val)
    outputs["log"] = log_file.replace(env.project_path(), "")
    env.dump(json.dumps(outputs, default=json_default_numpy), os.path.join(log_dir, ".outputs.json"))
    env.dump(json.dumps(opt_val, default=json_default_numpy), os.path.join(log_dir, ".metric"))
    return opt_val


def build_summary_json(logdir: str) -> str:
    """``{"combinations": [{"parameters": ..., "outputs": ...}, ...]}`` over all finished trials."""
    from .core.environment.singleton import EnvSing

    env = EnvSing.get_instance()
    combos = []
    for trial_dir in env.ls(logdir):
        if not env.isdir(trial_dir):
            continue
        ret, hp = os.path.join(trial_dir, ".outputs.json"), os.path.join(trial_dir, ".hparams.json")
        if env.exists(ret) and env.exists(hp):
            combos.append({"parameters": json.loads(env.load(hp)), "outputs": json.loads(env.load(ret))})
    return json.dumps({"combinations": combos}, default=json_default_numpy)


# ---------------------------------------------------------------------------------------------------
# experiment registration (ML_ID = <app_id>_<run_id>), reference maggy/util.py:216-243,275-290
# ---------------------------------------------------------------------------------------------------
def validate_ml_id(app_id, run_id):
    """Continue numbering if this process already ran experiments under the same app id."""
    ml_id = os.environ.get("ML_ID")
    if ml_id:
        prev_app, _, prev_run = ml_id.rpartition("_")
        try:
            if prev_app == str(app_id) and int(prev_run) >= run_id:
                run_id = int(prev_run) + 1
        except ValueError:
            pass
    return app_id, run_id


def set_ml_id(app_id, run_id):
    os.environ["HOME"] = os.environ.get("HOME", os.getcwd())
    os.environ["ML_ID"] = f"{app_id}_{run_id}"


def register_environment(app_id, run_id):
    """Create the experiment log dir, set ML_ID and register the tensorboard dir. Returns (app_id, run_id)."""
    from . import tensorboard
    from .core.environment.singleton import EnvSing

    app_id = str(app_id)
    app_id, run_id = validate_ml_id(app_id, run_id)
    set_ml_id(app_id, run_id)
    env = EnvSing.get_instance()
    env.create_experiment_dir(app_id, run_id)
    tensorboard._register(env.get_logdir(app_id, run_id))
    return app_id, run_id


def populate_experiment(config, app_id, run_id, exp_function):
    from .core.environment.singleton import EnvSing

    return EnvSing.get_instance().populate_experiment(config, app_id, run_id, exp_function)


def finalize_experiment(experiment_json, metric, app_id, run_id, state, duration, logdir, best_logdir,
                        optimization_key):
    from .core.environment.singleton import EnvSing

    EnvSing.get_instance().finalize_experiment(experiment_json, metric, app_id, run_id, state, duration,
                                               logdir, best_logdir, optimization_key)


def get_partition_attempt_id():
    """``(partition_id, attempt)`` of the calling worker process (reference: maggy/util.py:63-76 reads Spark's
    TaskContext). Here the native launcher passes both on the worker's command line; ``worker_main`` exports them as
    ``MAGGY_WORKER_ID`` / ``MAGGY_TASK_ATTEMPT``. In the driver process (no task context) this is ``(0, 0)``, like the
    reference without Spark."""
    return int(os.environ.get("MAGGY_WORKER_ID", "0")), int(os.environ.get("MAGGY_TASK_ATTEMPT", "0"))


def find_spark():
    """The SparkSession of the reference (maggy/util.py:246-253); this runtime never has one: workers are processes of
    the native per-GPU pool (core/workerpool.py), so the answer is always None, like the reference without Spark."""
    return None


def set_app_id(app_id):
    """Set the application id once (reference: maggy/util.py:339-347, used when the kernel is plain python)."""
    global APP_ID
    if APP_ID is None:
        APP_ID = app_id


def new_app_id() -> str:
    return f"application_{int(time.time() * 1000)}_0001"
