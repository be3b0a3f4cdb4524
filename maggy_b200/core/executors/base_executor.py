"""Executor for plain single runs (reference: maggy/core/executors/base_executor.py:20-41)."""
from __future__ import annotations

from .trial_executor import call_train_fn


def base_executor_fn(train_fn, config=None, reporter=None):
    """Return a callable that runs ``train_fn`` once in the calling process."""

    def _run(_=None):
        available = {"model": getattr(config, "model", None), "dataset": getattr(config, "dataset", None),
                     "hparams": getattr(config, "hparams", None) or {}, "reporter": reporter}
        return call_train_fn(train_fn, available)

    return _run


def base_executor_loop(spec, channel, worker_id, attempt):
    """Out-of-process variant (one worker, one run): used when the caller wants isolation."""
    channel.register()
    retval = base_executor_fn(spec["train_fn"], spec["config"])()
    channel.post_final("0" * 15 + "1", retval if isinstance(retval, (int, float)) else float("nan"), {"result": retval})
    channel.set_state(4)
    channel.post_exit()
