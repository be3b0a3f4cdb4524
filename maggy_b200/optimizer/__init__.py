from .abstractoptimizer import AbstractOptimizer
from .randomsearch import RandomSearch
from .asha import Asha
from .singlerun import SingleRun
from .gridsearch import GridSearch
from . import bayes

__all__ = ["AbstractOptimizer", "RandomSearch", "Asha", "SingleRun", "GridSearch", "bayes"]
