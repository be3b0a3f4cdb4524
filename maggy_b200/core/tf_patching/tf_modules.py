"""Strategy-scope model wrapper (reference: maggy/core/tf_patching/tf_modules.py:18-45).

The TensorFlow front-end is not rebuilt on B200 ("keep the config class for API parity"); asking for
the wrapper says so explicitly instead of failing on an import."""
from __future__ import annotations

from ..exceptions import NotSupportedError


def get_wrapped_model(model, strategy=None, is_chief: bool = False):
    raise NotSupportedError(
        "tf_patching", "TensorFlow", "maggy_b200 trains with PyTorch on sm_100a kernels; use TorchDistributedConfig")
