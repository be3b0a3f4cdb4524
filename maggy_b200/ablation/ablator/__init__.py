from .abstractablator import AbstractAblator
from .loco import LOCO

__all__ = ["AbstractAblator", "LOCO"]
