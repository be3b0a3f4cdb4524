"""Tracing (the reference has none beyond wall-clock job / trial durations).

Two layers, both free when switched off:

* **NVTX ranges** (`range("trial 3d1c…")`) around trials, training steps and user regions, so that an Nsight Systems /
  `ncu --nvtx` capture of a worker lines kernels up with the experiment's structure. No-ops without CUDA.
* a **host timeline** in Chrome trace-event format (`chrome://tracing`, Perfetto): every worker appends complete
  events (`ph: "X"`) for trials and, if the user wants, finer regions to `<log_dir>/trace_worker<k>.json`; the driver
  writes scheduling events (assign, final, stop) to `trace_driver.json`. Enabled with `MAGGY_TRACE=1`.
"""
from __future__ import annotations

import contextlib
import json
import os
import threading
import time

_lock = threading.Lock()
_sink = None          # open file object of this process' trace
_first = True
_pid_label = "proc"


def enabled() -> bool:
    return os.environ.get("MAGGY_TRACE", "0") not in ("", "0")


def start(path: str, label: str) -> None:
    """Open this process' trace file (idempotent)."""
    global _sink, _first, _pid_label
    with _lock:
        if _sink is not None or not enabled():
            return
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        _sink = open(path, "w")
        _sink.write("[\n")
        _first = True
        _pid_label = label


def _emit(ev: dict) -> None:
    global _first
    with _lock:
        if _sink is None:
            return
        _sink.write(("" if _first else ",\n") + json.dumps(ev))
        _first = False
        _sink.flush()


def instant(name: str, **args) -> None:
    if _sink is not None:
        _emit({"name": name, "ph": "i", "s": "p", "ts": time.time() * 1e6, "pid": _pid_label,
               "tid": threading.current_thread().name, "args": args})


def stop() -> None:
    global _sink
    with _lock:
        if _sink is not None:
            _sink.write("\n]\n")
            _sink.close()
            _sink = None


def _nvtx():
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.nvtx
    except Exception:  # noqa: BLE001
        pass
    return None


@contextlib.contextmanager
def range(name: str, **args):  # noqa: A001 - mirrors torch.cuda.nvtx.range
    """NVTX range + (when tracing is on) a complete event in the host timeline."""
    nv = _nvtx() if (enabled() or os.environ.get("MAGGY_NVTX", "0") not in ("", "0")) else None
    if nv is not None:
        nv.range_push(name)
    t0 = time.time()
    try:
        yield
    finally:
        if nv is not None:
            nv.range_pop()
        if _sink is not None:
            _emit({"name": name, "ph": "X", "ts": t0 * 1e6, "dur": (time.time() - t0) * 1e6, "pid": _pid_label,
                   "tid": threading.current_thread().name, "args": args})


def load(path: str) -> list:
    """Parse a trace file written by this module (also accepts one that was not closed cleanly)."""
    txt = open(path).read().rstrip()
    if not txt.endswith("]"):
        txt = txt.rstrip(",") + "\n]"
    return json.loads(txt)
