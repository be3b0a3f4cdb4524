"""Hyperparameter search space (API parity with reference maggy/searchspace.py:23-479).

Internally every hyperparameter is a small typed descriptor object (``_Double``, ``_Integer``,
``_Discrete``, ``_Categorical``) that knows how to validate, sample, encode and decode itself; the
``Searchspace`` is an insertion-ordered collection of those. The public surface (constructor kwargs,
``add``, attribute access to the feasible region, ``names/keys/values/items/get/to_dict/json``,
``get_random_parameter_values``, ``transform``/``inverse_transform``, ``dict_to_list``/``list_to_dict``)
behaves like the reference.
"""
from __future__ import annotations

import json
import random
from typing import Any, Dict, List

import numpy as np


class _Param:
    kind = ""

    def __init__(self, name: str, region):
        self.name = name
        self.region = region

    def sample(self, rng: random.Random):
        raise NotImplementedError

    def encode(self, value, normalize_categorical: bool):
        raise NotImplementedError

    def decode(self, value, normalize_categorical: bool):
        raise NotImplementedError


def _unit(lo, hi, x: float) -> float:
    return float(min(1.0, max(0.0, (float(x) - lo) / (hi - lo))))


class _Double(_Param):
    kind = "DOUBLE"

    def __init__(self, name, region):
        if any(type(b) not in (int, float) for b in region):
            raise ValueError(f"Hyperparameter boundaries for type DOUBLE need to be integer or float: {name}")
        super().__init__(name, region)

    def sample(self, rng):
        return rng.uniform(self.region[0], self.region[1])

    def encode(self, value, normalize_categorical):
        return _unit(self.region[0], self.region[1], value)

    def decode(self, value, normalize_categorical):
        lo, hi = self.region
        return float(value) * (hi - lo) + lo


class _Integer(_Param):
    kind = "INTEGER"

    def __init__(self, name, region):
        if any(type(b) is not int for b in region):
            raise ValueError(f"Hyperparameter boundaries for type INTEGER need to be integer: {name}")
        super().__init__(name, region)

    def sample(self, rng):
        return rng.randint(self.region[0], self.region[1])

    def encode(self, value, normalize_categorical):
        return _unit(self.region[0], self.region[1], int(value))

    def decode(self, value, normalize_categorical):
        lo, hi = self.region
        return int(np.round(float(value) * (hi - lo) + lo))


class _Discrete(_Param):
    kind = "DISCRETE"

    def sample(self, rng):
        return rng.choice(self.region)

    def encode(self, value, normalize_categorical):
        raise NotImplementedError("Not Implemented other types yet")

    decode = encode


class _Categorical(_Param):
    kind = "CATEGORICAL"

    def sample(self, rng):
        return rng.choice(self.region)

    def encode(self, value, normalize_categorical):
        code = self.region.index(value)
        if normalize_categorical:
            return _unit(0, len(self.region) - 1, code) if len(self.region) > 1 else 0.0
        return code

    def decode(self, value, normalize_categorical):
        if normalize_categorical:
            value = int(np.round(float(value) * (len(self.region) - 1)))
        return self.region[int(value)]


_KINDS = {c.kind: c for c in (_Double, _Integer, _Discrete, _Categorical)}


class Searchspace:
    """Insertion-ordered set of named hyperparameters.

    >>> sp = Searchspace(kernel=("INTEGER", [2, 8]), pool=("INTEGER", [2, 8]))
    >>> sp.add("dropout", ("DOUBLE", [0.01, 0.99]))
    >>> sp.kernel            # feasible region
    [2, 8]
    """

    DOUBLE = "DOUBLE"
    INTEGER = "INTEGER"
    DISCRETE = "DISCRETE"
    CATEGORICAL = "CATEGORICAL"

    def __init__(self, **kwargs):
        self._params: Dict[str, _Param] = {}
        self._rng = random
        for name, spec in kwargs.items():
            self.add(name, spec)

    # ------------------------------------------------------------------ construction
    def add(self, name: str, value) -> None:
        """Add ``name=(TYPE, feasible_region)``; DOUBLE/INTEGER regions are ``[lower, upper]``."""
        if getattr(self, name, None) is not None:
            raise ValueError(f"Hyperparameter name is reserved: {name}")
        if not isinstance(value, (tuple, list)):
            raise ValueError(f"Value is not an appropriate tuple: {name}")
        if len(value) != 2:
            raise ValueError(
                f"Hyperparameter tuple has to be of length two and format (type, list): {name}, {value}")
        kind, region = str(value[0]).upper(), value[1]
        if kind not in _KINDS:
            raise ValueError(
                f"Hyperparameter type is not of type DOUBLE, INTEGER, DISCRETE or CATEGORICAL: {name}")
        if len(region) == 0:
            raise ValueError(f"Hyperparameter feasible region list cannot be empty: {name}, {region}")
        if kind in (self.DOUBLE, self.INTEGER):
            assert len(region) == 2, (
                f"For DOUBLE or INTEGER type parameters, list can only contain upper and lower bounds: {name}, {region}")
        param = _KINDS[kind](name, region)
        if kind in (self.DOUBLE, self.INTEGER):
            assert region[0] < region[1], f"Lower bound {region[0]} must be less than upper bound {region[1]}: {name}"
        self._params[name] = param
        setattr(self, name, region)
        print(f"Hyperparameter added: {name}")

    def seed(self, seed) -> None:
        """Use a private, seeded RNG for sampling (reproducible experiments)."""
        self._rng = random.Random(seed)

    # ------------------------------------------------------------------ mapping-like access
    def names(self) -> Dict[str, str]:
        return {n: p.kind for n, p in self._params.items()}

    def keys(self) -> List[str]:
        return list(self._params)

    def values(self):
        return [(p.kind, p.region) for p in self._params.values()]

    def items(self):
        return iter(self)

    def get(self, name, default=None):
        return self._params[name].region if name in self._params else default

    def to_dict(self):
        return {n: (p.kind, p.region) for n, p in self._params.items()}

    def json(self) -> str:
        return json.dumps(self.to_dict(), sort_keys=True)

    __str__ = json

    def __contains__(self, name) -> bool:
        return name in self._params

    def __len__(self) -> int:
        return len(self._params)

    def __iter__(self):
        for n, p in self._params.items():
            yield {"name": n, "type": p.kind, "values": p.region}

    # ------------------------------------------------------------------ sampling
    def get_random_parameter_values(self, num: int) -> List[Dict[str, Any]]:
        if not isinstance(num, int):
            raise ValueError("num has to be an int")
        return [{n: p.sample(self._rng) for n, p in self._params.items()} for _ in range(num)]

    # ------------------------------------------------------------------ encoding for model-based optimizers
    def transform(self, hparams, normalize_categorical: bool = False):
        """Original representation (list ordered like the searchspace) -> numeric vector in [0,1]
        (categoricals: index, optionally min-max normalised)."""
        return [p.encode(v, normalize_categorical) for v, p in zip(hparams, self._params.values())]

    def inverse_transform(self, transformed_hparams, normalize_categorical: bool = False):
        return [p.decode(v, normalize_categorical) for v, p in zip(transformed_hparams, self._params.values())]

    @staticmethod
    def dict_to_list(hparams: dict) -> list:
        return list(hparams.values())

    def list_to_dict(self, hparams) -> dict:
        names = self.keys()
        if len(names) != len(hparams):
            raise ValueError("hparam_names and hparams have to have same length (and order!)")
        return dict(zip(names, hparams))
