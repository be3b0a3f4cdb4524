"""Distributed PyTorch training config (reference: maggy/config/torch_distributed.py:28-87).

``backend`` keeps the reference's vocabulary: ``"torch"`` selects the hand-written NVLink path of this
framework (``zero_lvl`` 0 = replicated optimizer state, 1/2 = fused reduce-scatter + sharded AdamW,
3 = fully sharded parameters with all-gather fused into the consuming GEMM); ``"deepspeed"`` is
accepted for API parity and maps its ``zero_optimization.stage`` onto the same levels (DeepSpeed itself
is not used).
"""
from __future__ import annotations

from typing import Optional

from .lagom import LagomConfig


class TorchDistributedConfig(LagomConfig):
    BACKENDS = ["torch", "deepspeed"]

    def __init__(self, module, dataset=None, hparams: Optional[dict] = None, backend: str = "torch",
                 mixed_precision: bool = False, zero_lvl: int = 0, deepspeed_config: Optional[dict] = None,
                 name: str = "torchDist", hb_interval: int = 1, description: str = "",
                 num_workers: Optional[int] = None):
        super().__init__(name, description, hb_interval)
        self.module = module
        self.dataset = dataset if dataset is not None else []
        self.hparams = hparams if hparams else {}
        if backend not in self.BACKENDS:
            raise ValueError(f"Backend {backend} not supported! Supported backends: {self.BACKENDS}")
        self.backend = backend
        self.mixed_precision = mixed_precision
        if zero_lvl not in (0, 1, 2, 3):
            raise ValueError(f"Zero level {zero_lvl} not supported! Pick one of 0, 1, 2, 3.")
        self.zero_lvl = zero_lvl
        if backend == "deepspeed":
            if not deepspeed_config:
                raise ValueError("DeepSpeed backend needs a deepspeed_config (a ZeRO stage at least)!")
            stage = int(deepspeed_config.get("zero_optimization", {}).get("stage", self.zero_lvl))
            self.zero_lvl = stage
        elif deepspeed_config:
            raise ValueError("deepspeed_config given although backend is not 'deepspeed'")
        self.deepspeed_config = deepspeed_config
        self.ds_config = deepspeed_config     # the attribute name the reference stores it under (torch_distributed.py:87)
        self.num_workers = num_workers
