"""TensorFlow distributed config (reference: maggy/config/tf_distributed.py:26-59).

``lagom(train_fn, TfDistributedConfig(...))`` starts one task per executor, derives chief / worker / evaluator
roles and every task's ``TF_CONFIG`` from the registrations and runs the train_fn under
``MultiWorkerMirroredStrategy`` (core/drivers/tf_distributed_training_driver.py, core/executors/tf_dist_executor.py).
TensorFlow itself is optional — this framework's compute path is PyTorch + sm_100a kernels; without it the
experiment fails with NotSupportedError and a pointer to :class:`TorchDistributedConfig`.
"""
from typing import Callable, Optional

from .lagom import LagomConfig


class TfDistributedConfig(LagomConfig):
    def __init__(self, model=None, dataset=None, process_data: Optional[Callable] = None,
                 mixed_precision: bool = False, name: str = "tfDist", hb_interval: int = 1,
                 description: str = "", hparams: Optional[dict] = None, num_workers: Optional[int] = None):
        super().__init__(name, description, hb_interval)
        self.model = model
        self.dataset = dataset
        self.process_data = process_data
        self.mixed_precision = mixed_precision
        self.hparams = hparams
        self.num_workers = num_workers   # default: one task per executor
