"""TensorFlow distributed training (reference:
maggy/core/experiment_driver/tf_distributed_training_driver.py:37-271).

The B200-native compute path of this framework is PyTorch + hand-written sm_100a kernels, so the
MultiWorkerMirroredStrategy executor is not rebuilt. The config class and this driver exist so that
code written against the reference fails with a clear message instead of an ImportError.
"""
from ..exceptions import NotSupportedError
from .base import Driver


class TfDistributedTrainingDriver(Driver):
    def __init__(self, config, app_id, run_id):
        try:
            import tensorflow  # noqa: F401
        except Exception as e:
            raise NotSupportedError(
                "framework", "tensorflow",
                " TensorFlow is not installed and TF distributed training is not part of the B200-native build; "
                "use TorchDistributedConfig (maggy_b200.parallel provides the fused NVLink data-parallel path).") from e
        raise NotSupportedError("framework", "tensorflow",
                                " TfDistributedConfig is kept for API compatibility only; use TorchDistributedConfig.")

    def _exp_startup_callback(self):
        pass

    def _exp_final_callback(self, job_end, exp_json):
        return {}

    def _register_msg_callbacks(self):
        pass
