"""Import-time stand-ins for optional third-party packages the reference imports at module scope but does not need
for a single-run (BaseConfig) experiment: any attribute resolves to a placeholder class."""
import importlib.abc
import importlib.machinery
import sys
import types

MISSING = ("skopt", "statsmodels", "deepspeed", "fairscale", "petastorm", "hops", "pydoop")


_STUBBED = set()


class _Placeholder:
    def __init__(self, *a, **k):
        raise RuntimeError("placeholder for a package that is not installed in this image")


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Placeholder,), {})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top in _STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        if top in MISSING:
            try:
                sys.meta_path.remove(self)
                real = importlib.util.find_spec(fullname.split(".")[0])
            except Exception:  # noqa: BLE001
                real = None
            finally:
                sys.meta_path.insert(0, self)
            if real is None:
                _STUBBED.add(top)
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


import importlib.util  # noqa: E402

sys.meta_path.insert(0, _Finder())
