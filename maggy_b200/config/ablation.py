"""Ablation-study config (reference: maggy/config/ablation.py:28-67)."""
from __future__ import annotations

from typing import Optional, Union

from ..ablation.ablationstudy import AblationStudy
from ..ablation.ablator import AbstractAblator
from .lagom import LagomConfig


class AblationConfig(LagomConfig):
    def __init__(self, ablation_study: AblationStudy, ablator: Union[str, AbstractAblator] = "loco",
                 direction: str = "max", name: str = "ablationStudy", description: str = "",
                 hb_interval: int = 1, model=None, dataset=None, optimization_key: str = "Metric",
                 gpus_per_trial: int = 1, num_workers: Optional[int] = None):
        super().__init__(name, description, hb_interval)
        if str(direction).lower() not in ("max", "min"):
            raise ValueError(f"direction must be 'max' or 'min', got {direction!r}")
        self.ablator = ablator
        self.ablation_study = ablation_study
        self.direction = str(direction).lower()
        self.model = model
        self.dataset = dataset
        self.optimization_key = optimization_key
        self.gpus_per_trial = gpus_per_trial
        self.num_workers = num_workers
