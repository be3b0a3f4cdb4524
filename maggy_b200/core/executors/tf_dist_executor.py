"""TensorFlow distributed executor (reference: maggy/core/executors/tf_dist_executor.py:35-480).

Not rebuilt: MultiWorkerMirroredStrategy's collectives are covered by the fused NVLink kernels of the torch path
(SURVEY.md §2.2 row "TF MultiWorkerMirroredStrategy"). The factory exists so that the driver can raise a precise error."""
from __future__ import annotations

from ..exceptions import NotSupportedError


def tf_dist_executor_fn(*args, **kwargs):
    raise NotSupportedError(
        "executor", "tf_dist_executor", "TensorFlow distributed training is not part of the B200 build; "
        "use TorchDistributedConfig (maggy_b200.core.executors.torch_dist_executor)")
