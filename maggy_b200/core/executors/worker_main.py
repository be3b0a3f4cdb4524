"""Entry point of every worker process started by the native launcher
(``python -m maggy_b200.core.executors.worker_main --segment ... --worker-id ... --spec ...``).

Plays the role of the cloudpickled ``executor_fn`` closure the reference ships to Spark tasks
(maggy/core/experiment_driver/spark_driver.py:136-145): it loads the experiment spec (train_fn,
config, paths), attaches to the control segment and runs the executor loop selected by the spec.
"""
from __future__ import annotations

import argparse
import os
import sys
import traceback


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--segment", required=True)
    ap.add_argument("--worker-id", type=int, required=True)
    ap.add_argument("--attempt", type=int, default=0)
    ap.add_argument("--spec", required=True)
    ap.add_argument("--gang-rank", type=int, default=0)
    ap.add_argument("--gang-size", type=int, default=1)
    ap.add_argument("--gang-port", type=int, default=0)
    args = ap.parse_args(argv)

    import cloudpickle

    with open(args.spec, "rb") as f:
        spec = cloudpickle.load(f)
    for p in spec.get("sys_path", []):
        if p not in sys.path:
            sys.path.append(p)
    os.environ["ML_ID"] = f"{spec['app_id']}_{spec['run_id']}"
    os.environ["MAGGY_WORKER_ID"] = str(args.worker_id)       # util.get_partition_attempt_id()
    os.environ["MAGGY_TASK_ATTEMPT"] = str(args.attempt)

    from ..environment.singleton import EnvSing
    from .. import rpc

    env = EnvSing.get_instance()
    if spec.get("env_log_dir"):
        env.log_dir = spec["env_log_dir"]

    gang = None
    if args.gang_size > 1:
        from .gang import Gang

        gang = Gang(args.gang_rank, args.gang_size, args.gang_port)

    return run_executor(spec, args.segment, args.worker_id, args.attempt, gang, args.gang_rank)


def run_executor(spec: dict, segment: str, worker_id: int, attempt: int = 0, gang=None, gang_rank: int = 0) -> int:
    """Attach to the control segment as worker ``worker_id`` and run the executor loop ``spec["kind"]`` selects — the
    body of the closure the reference ships to a Spark task (spark_driver.py:136-145). Also what the drivers'
    ``_patching_fn`` and the ``*_executor_fn`` factories return, so the same loop can be run in a process that was not
    started by the native launcher."""
    from .. import rpc

    channel = None
    if gang_rank == 0:
        channel = rpc.WorkerChannel(segment, worker_id, attempt=attempt)
    kind = spec["kind"]
    try:
        if kind == "trial":
            from .trial_executor import trial_executor_loop

            trial_executor_loop(spec, channel, worker_id, attempt, gang)
        elif kind == "torch_dist":
            from .torch_dist_executor import torch_dist_executor_loop

            torch_dist_executor_loop(spec, channel, worker_id, attempt)
        elif kind == "tf_dist":
            from .tf_dist_executor import tf_dist_executor_loop

            tf_dist_executor_loop(spec, channel, worker_id, attempt)
        elif kind == "base":
            from .base_executor import base_executor_loop

            base_executor_loop(spec, channel, worker_id, attempt)
        else:
            raise ValueError(f"unknown executor kind {kind}")
    except Exception:
        text = traceback.format_exc()
        sys.stderr.write(text)
        if channel is not None:
            try:
                channel.post_error(None, text)
            except Exception:
                pass
        return 1
    finally:
        if gang is not None:
            gang.close()
        if channel is not None:
            channel.close()
    return 0


def make_executor_fn(kind: str, train_fn, config, experiment_type, app_id, run_id, server_addr, hb_interval, secret,
                     optimization_key, log_dir):
    """``executor_fn(partition_id)`` with the argument list of the reference's executor factories
    (trial_executor.py:35-60, torch_dist_executor.py:63-90, tf_dist_executor.py:35-60). ``server_addr`` is the control
    segment's name; ``secret`` is unused (the segment name is the credential)."""
    from ..environment.singleton import EnvSing

    spec = {"kind": kind, "train_fn": train_fn, "config": config, "log_dir": log_dir,
            "env_log_dir": EnvSing.get_instance().log_dir, "app_id": app_id, "run_id": run_id, "hb_interval": hb_interval,
            "experiment_type": experiment_type, "optimization_key": optimization_key, "sys_path": []}

    def _wrapper_fun(partition_id, attempt: int = 0):
        return run_executor(spec, server_addr, int(partition_id), attempt)

    return _wrapper_fun


if __name__ == "__main__":
    sys.exit(main())
