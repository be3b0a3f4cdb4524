from .arena import FlatParams, FusedAdamW
from .trainer import TrainStep, build_gpt2_trainer

__all__ = ["FlatParams", "FusedAdamW", "TrainStep", "build_gpt2_trainer", "DataParallel"]


def __getattr__(name):
    if name == "DataParallel":
        from .data_parallel import DataParallel

        return DataParallel
    raise AttributeError(name)
