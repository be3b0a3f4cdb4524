from .lagom import LagomConfig
from .base_config import BaseConfig
from .ablation import AblationConfig
from .hyperparameter_optimization import HyperparameterOptConfig, OptimizationConfig
from .torch_distributed import TorchDistributedConfig
from .tf_distributed import TfDistributedConfig

__all__ = ["LagomConfig", "BaseConfig", "AblationConfig", "HyperparameterOptConfig", "OptimizationConfig",
           "TfDistributedConfig", "TorchDistributedConfig"]
