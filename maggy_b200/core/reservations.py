"""Worker reservations (reference: maggy/core/rpc.py:45-123 ``Reservations``, :126-202 ``MirroredReservations``).

In the reference these dictionaries ARE the registry (filled by REG messages over TCP). Here the source of truth is the
per-worker slot header in the shared-memory control segment (``csrc/ctrl/ring.cpp``: state, pid, heartbeat, assigned
trial); the classes below are the thread-safe Python view the drivers and the tests use, kept in sync by
:class:`maggy_b200.core.rpc.DriverChannel` as REG / assignment events pass through it.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional


class Reservations:
    """``{partition_id: {"host_port", "task_attempt", "trial_id", "num_executors", "pid"}}`` with the reference API."""

    def __init__(self, required: int):
        self.required = required
        self.lock = threading.RLock()
        self.reservations: Dict[int, dict] = {}
        self.check_done = False

    def add(self, meta: dict) -> None:
        with self.lock:
            self.reservations[int(meta["partition_id"])] = {
                "host_port": meta.get("host_port"), "task_attempt": meta.get("task_attempt", 0),
                "trial_id": meta.get("trial_id"), "num_executors": self.required, "pid": meta.get("pid")}
            if self.remaining() == 0:
                self.check_done = True

    def done(self) -> bool:
        with self.lock:
            return self.check_done

    def get(self) -> Dict[int, dict]:
        with self.lock:
            return {k: dict(v) for k, v in self.reservations.items()}

    def remaining(self) -> int:
        with self.lock:
            return self.required - len(self.reservations)

    def get_assigned_trial(self, partition_id: int) -> Optional[str]:
        with self.lock:
            r = self.reservations.get(int(partition_id))
            return None if r is None else r.get("trial_id")

    def assign_trial(self, partition_id: int, trial_id: Optional[str]) -> None:
        with self.lock:
            r = self.reservations.get(int(partition_id))
            if r is not None:
                r["trial_id"] = trial_id


class MirroredReservations(Reservations):
    """Adds the cluster specification the reference's TensorFlow server derives from the registrations: the first
    registered worker is ``chief``, the last one ``evaluator`` (if requested), the rest ``worker``."""

    def __init__(self, required: int, evaluator: bool = False):
        super().__init__(required)
        self.evaluator = evaluator

    def add(self, meta: dict) -> None:
        super().add(meta)
        with self.lock:
            r = self.reservations[int(meta["partition_id"])]
            r["rank"] = len(self.reservations) - 1

    def get_spec(self) -> Dict[str, List[str]]:
        with self.lock:
            ordered = [r for _, r in sorted(self.reservations.items(), key=lambda kv: kv[1].get("rank", kv[0]))]
        hosts = [r["host_port"] for r in ordered]
        spec: Dict[str, List[str]] = {"chief": hosts[:1]}
        rest = hosts[1:]
        if self.evaluator and rest:
            spec["evaluator"] = rest[-1:]
            rest = rest[:-1]
        if rest:
            spec["worker"] = rest
        return spec
