from .abstractpruner import AbstractPruner
from .hyperband import Hyperband

__all__ = ["AbstractPruner", "Hyperband"]
