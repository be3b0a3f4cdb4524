"""Runtime mode flags (reference: maggy/core/config.py:17-47).

The reference probes for PySpark and parses the TensorFlow major version at import time. Here the substrate is fixed —
a native per-GPU worker pool and CUDA on sm_100a — so the probes answer accordingly and stay import-light."""
from __future__ import annotations

import importlib.util
import os

SPARK_AVAILABLE = False           # the Spark executor layer is replaced by csrc/runtime/launcher.cpp
MODE = None
TF_VERSION = None


def initialize() -> None:
    """Detect the optional front-ends once (TensorFlow is only ever *detected*, never imported here)."""
    global MODE, TF_VERSION
    MODE = "native-worker-pool"
    if TF_VERSION is None and importlib.util.find_spec("tensorflow") is not None:
        try:
            from importlib.metadata import version

            TF_VERSION = int(version("tensorflow").split(".")[0])
        except Exception:  # noqa: BLE001
            TF_VERSION = None


def is_spark_available() -> bool:
    return SPARK_AVAILABLE


def is_cuda_available() -> bool:
    if os.environ.get("CUDA_VISIBLE_DEVICES") == "":
        return False
    import torch

    return torch.cuda.is_available()
