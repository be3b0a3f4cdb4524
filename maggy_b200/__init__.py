"""maggy_b200 — a B200-native experiment orchestrator with the capabilities of logicalclocks/maggy.

    from maggy_b200 import experiment, Searchspace
    result = experiment.lagom(train_fn, config)

Public surface mirrors the reference (maggy/__init__.py:17-21): ``experiment`` (sub-package with
``lagom``), ``Searchspace``; plus ``config``, ``optimizer``, ``pruner``, ``earlystop``, ``ablation``,
``callbacks``, ``tensorboard``, ``Trial``. The compute side lives in ``maggy_b200.ops`` (sm_100a kernels),
``maggy_b200.parallel`` (fused NVLink data parallelism) and ``maggy_b200.models``.

Sub-packages are imported lazily so that worker processes (which only need the executor loop) start fast.
"""
import importlib

from .version import __version__
from .searchspace import Searchspace

__all__ = ["Searchspace", "Trial", "experiment", "config", "optimizer", "pruner", "earlystop", "ablation",
           "callbacks", "tensorboard", "ops", "parallel", "models", "__version__"]

_LAZY = {"experiment", "config", "optimizer", "pruner", "earlystop", "ablation", "callbacks", "tensorboard",
         "ops", "parallel", "models", "trial", "util", "core", "constants"}


def __getattr__(name):
    if name == "Trial":
        from .trial import Trial

        return Trial
    if name in _LAZY:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
