"""``experiment.lagom`` — the single public entry point (reference: maggy/experiment/experiment.py:21-45,
experiment_pyspark.py:43-183, experiment_python.py:48-197).

*lagom* (Swedish): just the right amount. ``lagom(train_fn, config)`` dispatches on the type of
``config`` to the matching driver, runs the experiment and returns its result dict.
"""
from __future__ import annotations

import atexit
import time
from functools import singledispatch
from typing import Callable, Optional

from .. import util
from ..config import (AblationConfig, BaseConfig, HyperparameterOptConfig, LagomConfig, TfDistributedConfig,
                      TorchDistributedConfig)
from ..core.environment.singleton import EnvSing

APP_ID = None
RUNNING = False
RUN_ID = 1
EXPERIMENT_JSON = {}


def lagom(train_fn: Callable, config: Optional[LagomConfig] = None) -> dict:
    """Launch a maggy experiment.

    :param train_fn: the user's training function; it may declare any of the arguments
        ``model / module, dataset, hparams, reporter`` and returns a number or a dict with the metric.
    :param config: ``BaseConfig`` (default: run once), ``HyperparameterOptConfig``, ``AblationConfig``,
        ``TorchDistributedConfig`` (or ``TfDistributedConfig``).
    :returns: the experiment result dict.
    """
    global APP_ID, RUNNING, RUN_ID, EXPERIMENT_JSON
    if config is None:
        config = BaseConfig(name="maggyExperiment", description="experiment without config object", hb_interval=1)
    job_start = time.time()
    if RUNNING:
        raise RuntimeError("An experiment is currently running.")
    driver = None
    try:
        RUNNING = True
        if APP_ID is None:
            APP_ID = util.new_app_id()
        APP_ID, RUN_ID = util.register_environment(APP_ID, RUN_ID)
        EnvSing.get_instance().set_ml_id(APP_ID, RUN_ID)
        driver = lagom_driver(config, APP_ID, RUN_ID)
        return driver.run_experiment(train_fn, config)
    except Exception:
        _exception_handler(util.seconds_to_milliseconds(time.time() - job_start))
        raise
    finally:
        RUN_ID += 1
        RUNNING = False


@singledispatch
def lagom_driver(config, app_id: str, run_id: int):
    raise TypeError(f"Invalid config type! Config is expected to be of type {LagomConfig}, but is of type "
                    f"{type(config)}")


@lagom_driver.register(BaseConfig)
def _(config: BaseConfig, app_id, run_id):
    from ..core.drivers.base_driver import BaseDriver

    return BaseDriver(config, app_id, run_id)


@lagom_driver.register(HyperparameterOptConfig)
def _(config: HyperparameterOptConfig, app_id, run_id):
    from ..core.drivers.optimization_driver import HyperparameterOptDriver

    return HyperparameterOptDriver(config, app_id, run_id)


@lagom_driver.register(AblationConfig)
def _(config: AblationConfig, app_id, run_id):
    from ..core.drivers.ablation_driver import AblationDriver

    return AblationDriver(config, app_id, run_id)


@lagom_driver.register(TorchDistributedConfig)
def _(config: TorchDistributedConfig, app_id, run_id):
    from ..core.drivers.torch_distributed_training_driver import TorchDistributedTrainingDriver

    return TorchDistributedTrainingDriver(config, app_id, run_id)


@lagom_driver.register(TfDistributedConfig)
def _(config: TfDistributedConfig, app_id, run_id):
    from ..core.drivers.tf_distributed_training_driver import TfDistributedTrainingDriver

    return TfDistributedTrainingDriver(config, app_id, run_id)


def _exception_handler(duration: int) -> None:
    """Mark the experiment FAILED in its metadata (reference experiment_pyspark.py:149-165)."""
    try:
        global EXPERIMENT_JSON
        if RUNNING and EXPERIMENT_JSON is not None:
            EXPERIMENT_JSON["state"] = "FAILED"
            EXPERIMENT_JSON["duration"] = duration
            EnvSing.get_instance().attach_experiment_xattr(f"{APP_ID}_{RUN_ID}", EXPERIMENT_JSON, "FULL_UPDATE")
    except Exception as err:  # pragma: no cover
        util.log(err)


def _exit_handler() -> None:
    """Mark a still-running experiment KILLED at interpreter exit (reference experiment_pyspark.py:168-183)."""
    try:
        global EXPERIMENT_JSON
        if RUNNING and EXPERIMENT_JSON is not None:
            EXPERIMENT_JSON["status"] = "KILLED"
            EnvSing.get_instance().attach_experiment_xattr(f"{APP_ID}_{RUN_ID}", EXPERIMENT_JSON, "FULL_UPDATE")
    except Exception as err:  # pragma: no cover
        util.log(err)


atexit.register(_exit_handler)
