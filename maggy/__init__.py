"""``maggy`` import alias: user code written against logicalclocks/maggy keeps its imports.

``from maggy import experiment, Searchspace`` (reference: maggy/__init__.py:17-21),
``from maggy.optimizer import Asha``, ``from maggy.experiment_config import OptimizationConfig``,
``from maggy.ablation import AblationStudy`` ... all resolve to the corresponding ``maggy_b200`` module —
the SAME module object, so ``isinstance`` checks and singletons are shared between the two spellings.
"""
import importlib
import importlib.abc
import importlib.machinery
import sys

import maggy_b200 as _impl

_PREFIX = __name__ + "."
# legacy module names used by the reference's notebooks / docs
_RENAMED = {"maggy.experiment_config": "maggy_b200.config"}


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PREFIX):
            return None
        real = _RENAMED.get(fullname, "maggy_b200." + fullname[len(_PREFIX):])
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        spec = importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        spec._maggy_real = real
        return spec

    def create_module(self, spec):
        return importlib.import_module(spec._maggy_real)

    def exec_module(self, module):  # the real module is already initialised
        return None


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

__version__ = _impl.__version__
__all__ = list(getattr(_impl, "__all__", []))
__path__ = []  # submodules come from the alias finder only


def __getattr__(name):
    try:
        return getattr(_impl, name)
    except AttributeError:
        try:
            return importlib.import_module(_PREFIX + name)
        except ImportError:
            raise AttributeError(f"module 'maggy' has no attribute {name!r}") from None
