from .abstractearlystop import AbstractEarlyStop
from .medianrule import MedianStoppingRule
from .nostop import NoStoppingRule

__all__ = ["AbstractEarlyStop", "MedianStoppingRule", "NoStoppingRule"]
