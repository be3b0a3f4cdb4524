"""Trial: one evaluation of a hyperparameter combination (reference: maggy/trial.py:24-176).

The trial id is the first 16 hex chars (64 bits) of the md5 of the canonical JSON of the params, so
ids are stable across processes and identical to the reference's (golden value in tests:
``{"param1": 5, "param2": "ada"}`` -> ``3d1cc9fdb1d4d001``). The 64-bit integer form of the id is
what travels through the control rings (``MgRecord.trial_id``).
"""
from __future__ import annotations

import hashlib
import json
import threading
from typing import Optional

from . import util


class Trial:
    PENDING = "PENDING"
    SCHEDULED = "SCHEDULED"
    RUNNING = "RUNNING"
    ERROR = "ERROR"
    FINALIZED = "FINALIZED"

    _SERIALIZED = ("trial_type", "trial_id", "params", "status", "early_stop", "final_metric",
                   "metric_history", "step_history", "metric_dict", "duration", "info_dict")

    def __init__(self, params: dict, trial_type: str = "optimization", info_dict: Optional[dict] = None):
        self.trial_type = trial_type
        if trial_type == "ablation":
            # callables (dataset/model generators) are not hashable content: id = what was ablated
            id_src = {"ablated_feature": params.get("ablated_feature", None),
                      "ablated_layer": params.get("ablated_layer", None)}
        else:
            id_src = params
        self.trial_id = Trial._generate_id(id_src)
        self.params = params
        self.status = Trial.PENDING
        self.early_stop = False
        self.final_metric = None
        self.metric_history = []
        self.step_history = []
        self.metric_dict = {}
        self.start = None
        self.duration = None
        self.lock = threading.RLock()
        self.info_dict = {} if info_dict is None else info_dict

    # ------------------------------------------------------------------ ids
    @classmethod
    def _generate_id(cls, params) -> str:
        if not isinstance(params, dict):
            raise ValueError("Hyperparameters need to be a dictionary.")
        if not all(isinstance(k, str) for k in params):
            raise ValueError("All hyperparameter names have to be strings.")
        blob = json.dumps(params, sort_keys=True, default=util.json_default_numpy).encode("utf-8")
        return hashlib.md5(blob).hexdigest()[:16]

    @property
    def id64(self) -> int:
        """The id as the unsigned 64-bit integer carried in ring records."""
        return int(self.trial_id, 16)

    @staticmethod
    def id_from_int(v: int) -> str:
        return f"{v:016x}"

    # ------------------------------------------------------------------ early stop flag
    def get_early_stop(self) -> bool:
        with self.lock:
            return self.early_stop

    def set_early_stop(self) -> None:
        with self.lock:
            self.early_stop = True

    # ------------------------------------------------------------------ metric stream
    def append_metric(self, metric_data: dict):
        """Record ``{"value": v, "step": s}``; returns the step if it was new, else ``None``."""
        with self.lock:
            step, value = metric_data["step"], metric_data["value"]
            if value is None or step in self.metric_dict:
                return None
            self.metric_dict[step] = value
            self.metric_history.append(value)
            self.step_history.append(step)
            return step

    # ------------------------------------------------------------------ (de)serialisation
    def to_dict(self) -> dict:
        d = {"__class__": type(self).__name__}
        with self.lock:
            for k in self._SERIALIZED:
                d[k] = getattr(self, k)
        return d

    def to_json(self) -> str:
        return json.dumps(self.to_dict(), default=util.json_default_numpy)

    @classmethod
    def from_json(cls, json_str: str) -> "Trial":
        d = json.loads(json_str)
        if d.get("__class__") != "Trial":
            raise ValueError("json_str is not a Trial object.")
        inst = cls(d.get("params") or {}, trial_type=d.get("trial_type", "optimization"),
                   info_dict=d.get("info_dict"))
        inst.trial_id = d["trial_id"]
        inst.status = d["status"]
        inst.early_stop = d.get("early_stop", False)
        inst.final_metric = d["final_metric"]
        inst.metric_history = d.get("metric_history", [])
        inst.step_history = d.get("step_history", [])
        # JSON object keys are strings: restore integer steps where possible
        md = {}
        for k, v in (d.get("metric_dict") or {}).items():
            try:
                md[int(k)] = v
            except (TypeError, ValueError):
                md[k] = v
        inst.metric_dict = md
        inst.duration = d.get("duration")
        return inst

    def __repr__(self):
        return f"Trial({self.trial_id}, {self.status}, params={self.params})"
