from .base import BaseAsyncBO
from .gp import GP
from .tpe import TPE

__all__ = ["BaseAsyncBO", "TPE", "GP"]
