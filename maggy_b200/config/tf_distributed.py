"""TensorFlow distributed config (reference: maggy/config/tf_distributed.py:26-59).

Kept for API parity. This framework's compute path is PyTorch + sm_100a kernels; ``lagom`` with a
``TfDistributedConfig`` runs only if TensorFlow is importable and otherwise raises NotSupportedError
with a pointer to :class:`TorchDistributedConfig`.
"""
from typing import Callable, Optional

from .lagom import LagomConfig


class TfDistributedConfig(LagomConfig):
    def __init__(self, model=None, dataset=None, process_data: Optional[Callable] = None,
                 mixed_precision: bool = False, name: str = "tfDist", hb_interval: int = 1,
                 description: str = "", hparams: Optional[dict] = None):
        super().__init__(name, description, hb_interval)
        self.model = model
        self.dataset = dataset
        self.process_data = process_data
        self.mixed_precision = mixed_precision
        self.hparams = hparams
