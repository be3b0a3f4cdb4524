"""Worker of a distributed PyTorch training run (reference: maggy/core/executors/torch_dist_executor.py:63-422).

Flow: register -> receive (rank, world_size, master port) with the assignment -> environment variables ->
``init_process_group`` (NCCL on GPUs, gloo on CPU; 60 s timeout like the reference) -> seeds -> wrap the
user's module class (patching/modules.py) -> scoped optimizer / DataLoader patching -> ``train_fn(module=,
dataset=, hparams=, reporter=)`` -> validate the return value -> barrier -> FINAL.
"""
from __future__ import annotations

import builtins
import datetime
import os
import random
import time
import traceback

import numpy as np

from ... import util
from .. import rpc
from ..reporter import Reporter
from .trial_executor import _start_heartbeat, call_train_fn


def _init_cluster(rank, world, addr, port, timeout_s=60):
    import torch
    import torch.distributed as dist

    local = int(os.environ.get("MAGGY_LOCAL_RANK", "0"))
    os.environ.update({"MASTER_ADDR": addr, "MASTER_PORT": str(port), "WORLD_SIZE": str(world), "RANK": str(rank),
                       "LOCAL_RANK": str(local)})
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)  # all ranks see all GPUs (peer-mapped symmetric memory), each drives its own
    if world > 1 or True:
        dist.init_process_group("nccl" if use_cuda else "gloo", rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=timeout_s))
    torch.manual_seed(0)
    np.random.seed(0)
    random.seed(0)
    if use_cuda:
        torch.cuda.manual_seed_all(0)
    return dist


def _wrap_module_dispatcher(config):
    """Pick the class factory for the configured backend / ZeRO level (reference :288-357)."""
    from ..patching import modules as pm

    module = config.module
    if module is None:
        return None
    if getattr(config, "backend", "torch") == "deepspeed" and config.deepspeed_config:
        # an engine with DeepSpeed's API (backward / step / loss scaling) on the fused data plane
        wrap_ds = lambda m: pm.get_maggy_deepspeed_wrapper(m, config.deepspeed_config)  # noqa: E731
        if isinstance(module, (list, tuple)):
            return type(module)(wrap_ds(m) for m in module)
        return wrap_ds(module)
    if config.zero_lvl >= 3:
        wrap = lambda m: pm.get_maggy_fsdp_wrapper(m, config.mixed_precision)  # noqa: E731
    else:
        wrap = pm.get_maggy_ddp_wrapper
    if isinstance(module, (list, tuple)):
        return type(module)(wrap(m) for m in module)
    return wrap(module)


def torch_dist_executor_loop(spec, channel, worker_id, attempt):
    import torch

    from ..patching import dataloader as pdl
    from ..patching.optim import patched_optimizers

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
    dist = None
    clean_exit = False
    saved_loader = torch.utils.data.DataLoader
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
        dist = _init_cluster(payload["rank"], payload["world_size"], payload["master_addr"], payload["master_port"])
        reporter.log(f"Rank {payload['rank']}/{payload['world_size']} joined the process group", False)
        module = _wrap_module_dispatcher(config)
        torch.utils.data.DataLoader = pdl.MaggyDataLoader  # scoped: restored in finally
        t0 = time.time()
        with patched_optimizers(config.zero_lvl):
            retval = call_train_fn(spec["train_fn"], {"module": module, "model": module, "dataset": config.dataset,
                                                      "hparams": config.hparams, "reporter": reporter})
        metric = util.handle_return_val(retval, log_dir, "Metric", os.path.join(log_dir, "output.log"))
        dist.barrier()
        _, _, logs = reporter.get_data()
        channel.post_final(trial_id, metric, {"logs": logs, "wall_s": time.time() - t0})
        clean_exit = True
    except Exception:
        text = traceback.format_exc()
        reporter.log(text, False)
        channel.post_error(None, text)
        raise
    finally:
        torch.utils.data.DataLoader = saved_loader
        hb_stop.set()
        builtins.print = original_print
        try:
            if dist is not None and dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass
        try:
            channel.set_state(rpc.W_DONE if clean_exit else rpc.W_DEAD)
            if clean_exit:
                channel.post_exit()
        except Exception:
            pass
        reporter.close_logger()



def torch_dist_executor_fn(train_fn, config, experiment_type, app_id, run_id, server_addr, hb_interval, secret, optimization_key, log_dir):
    """The reference's executor factory (same argument list): ``executor_fn(partition_id)`` that attaches to the control
    segment ``server_addr`` and runs this module's loop — see ``worker_main.make_executor_fn``."""
    from .worker_main import make_executor_fn

    return make_executor_fn("torch_dist", train_fn, config, experiment_type, app_id, run_id, server_addr, hb_interval, secret,
                            optimization_key, log_dir)
