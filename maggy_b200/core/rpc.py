"""Control-plane endpoints over the shared-memory rings (replaces reference maggy/core/rpc.py).

The reference's ``Server``/``Client`` pair speaks length-prefixed cloudpickle over TCP with a 1 s
heartbeat and a 1 s GET poll (maggy/core/rpc.py:205-257, 716-748). Here both ends are thin ctypes
wrappers over ``libmaggy_b200_host.so`` (csrc/ctrl/ring.cpp):

========================  =====================================================================
reference message         here
========================  =====================================================================
REG                       ``WorkerChannel.register`` -> REG record
METRIC (heartbeat)        ``WorkerChannel.post_metric`` (host) / ``post_metric_device`` (GPU kernel K8)
STOP (heartbeat reply)    per-worker stop word, written by the native policy engine / K9 / driver
FINAL                     ``WorkerChannel.post_final`` -> FINAL record + pickled blob
GET / TRIAL               ``DriverChannel.assign`` writes the slot; ``WorkerChannel.wait_assign`` futex-waits
GSTOP                     ``DriverChannel.gstop``
LOG                       ``WorkerChannel.post_log`` -> LOG record + text blob
BLACK                     ``DriverChannel.dead_workers`` (heartbeat timeout) + re-REG with attempt > 0
QUERY / EXEC_CONFIG       not needed: rendezvous facts travel in the assign payload of gang trials
secret (auth)             the segment is created 0600 under a random name; only same-user processes
                          that were handed the name can attach
========================  =====================================================================
"""
from __future__ import annotations

import ctypes
import os
import secrets
import time
from typing import List, Optional

import cloudpickle

from ..ops import _lib

REC_REG, REC_METRIC, REC_FINAL, REC_LOG, REC_IDLE, REC_ERROR, REC_HEARTBEAT, REC_EXIT = 1, 2, 3, 4, 5, 6, 7, 8
REC_NAMES = {1: "REG", 2: "METRIC", 3: "FINAL", 4: "LOG", 5: "IDLE", 6: "ERROR", 7: "HEARTBEAT", 8: "EXIT"}
W_EMPTY, W_REGISTERED, W_IDLE, W_RUNNING, W_DONE, W_DEAD = 0, 1, 2, 3, 4, 5


class MgRecord(ctypes.Structure):
    _fields_ = [("seq", ctypes.c_uint64), ("type", ctypes.c_uint16), ("worker", ctypes.c_uint16),
                ("aux", ctypes.c_uint32), ("trial_id", ctypes.c_uint64), ("step", ctypes.c_int64),
                ("value", ctypes.c_double), ("t_ns", ctypes.c_uint64), ("reserved", ctypes.c_uint64 * 2)]


assert ctypes.sizeof(MgRecord) == 64

_configured = False


def _host():
    global _configured
    lib = _lib.host()
    if _configured:
        return lib
    P, u32, u64, i64, i32, dbl = (ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64,
                                  ctypes.c_int, ctypes.c_double)

    def sig(name, argtypes, restype=i32):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argtypes, restype

    sig("mgr_create", [ctypes.c_char_p, u32, u32, u32], P)
    sig("mgr_attach", [ctypes.c_char_p], P)
    sig("mgr_close", [P, i32], None)
    sig("mgr_base", [P], P)
    sig("mgr_bytes", [P], u64)
    sig("mgr_num_workers", [P], u32)
    sig("mgr_payload_bytes", [P], u32)
    sig("mgr_set_mailbox", [P, ctypes.c_char_p, u32, u64], None)
    sig("mgr_clear_mailbox", [P], None)
    sig("mgr_get_mailbox", [P, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32)], u32)
    sig("mgr_now_ns", [], u64)
    sig("mgr_off_dev_ring", [P, u32], u64)
    sig("mgr_off_ctl", [P, u32], u64)
    sig("mgr_ring_slots", [P], u32)
    sig("mgr_post", [P, u32, u32, u64, i64, dbl, u32])
    sig("mgr_post_blob", [P, u32, u32, u64, i64, dbl, ctypes.c_char_p, u32, i64])
    sig("mgr_register", [P, u32, u32, u32])
    sig("mgr_set_state", [P, u32, u32], None)
    sig("mgr_heartbeat", [P, u32], None)
    sig("mgr_wait_assign", [P, u32, i64, ctypes.POINTER(u64), P, u32, ctypes.POINTER(u32)])
    sig("mgr_check_stop", [P, u32, u64])
    sig("mgr_gstop_set", [P])
    sig("mgr_assign", [P, u32, u64, ctypes.c_char_p, u32])
    sig("mgr_assign_tagged", [P, u32, u64, ctypes.c_char_p, u32, u32])
    sig("mgr_stage", [P, u32, u64, ctypes.c_char_p, u32, u32])
    sig("mgr_staged_state", [P, u32, ctypes.POINTER(u64)])
    sig("mgr_publish_staged", [P, u32])
    sig("mgr_decision", [P, u32, ctypes.POINTER(u32), ctypes.POINTER(u32), ctypes.POINTER(u64), ctypes.POINTER(u64)], u32)
    sig("mgr_asha_create", [u32, u32, u32, ctypes.c_int], P)
    sig("mgr_asha_free", [P], None)
    sig("mgr_asha_final", [P, u32, u64, ctypes.c_double, ctypes.POINTER(u64), ctypes.POINTER(u32)], u32)
    sig("mgr_stop_trial", [P, u32, u64], None)
    sig("mgr_gstop", [P], None)
    sig("mgr_read_blob", [P, u32, P, u32])
    sig("mgr_worker_heartbeat", [P, u32], u64)
    sig("mgr_worker_state", [P, u32], u32)
    sig("mgr_worker_pid", [P, u32], u32)
    sig("mgr_dev_dropped", [P, u32], u64)
    sig("mgr_stop_time_ns", [P, u32], u64)
    sig("mgr_policy_config", [P, u32, i32, u32, u32], None)
    sig("mgr_policy_stops", [P], u64)
    sig("mgr_start_pump", [P], None)
    sig("mgr_stop_pump", [P], None)
    sig("mgr_next", [P, ctypes.POINTER(MgRecord), i32, i64])
    sig("mgr_poll", [P, ctypes.POINTER(MgRecord), i32])
    _configured = True
    return lib


def now_ns() -> int:
    return int(_host().mgr_now_ns())


def new_segment_name() -> str:
    return f"/maggy_b200_{os.getpid()}_{secrets.token_hex(8)}"


class DriverChannel:
    """Driver end: owns the segment, drains records, assigns trials, issues stops."""

    def __init__(self, num_workers: int, ring_slots: int = 1024, payload_bytes: int = 1 << 20,
                 name: Optional[str] = None, pump: bool = True):
        self.lib = _host()
        self.name = name or new_segment_name()
        self.num_workers = num_workers
        self.payload_bytes = payload_bytes
        self.h = self.lib.mgr_create(self.name.encode(), num_workers, ring_slots, payload_bytes)
        if not self.h:
            raise OSError(f"could not create control segment {self.name}")
        self._buf = (MgRecord * 256)()
        self._blob = ctypes.create_string_buffer(int(self.lib.mgr_bytes(self.h)) and payload_bytes + 4096)
        self.pump = pump
        # Python view of the worker slots with the reference's Reservations API (core/reservations.py)
        from .reservations import Reservations

        self.reservations = Reservations(num_workers)
        if pump:
            self.lib.mgr_start_pump(self.h)

    # -- receive ---------------------------------------------------------------------------------
    def next(self, timeout_s: float = 0.05) -> List[dict]:
        """Blocking fetch (releases the GIL). Each message is a dict mirroring the reference's RPC
        messages: ``{"type", "partition_id", "trial_id", "data", "step", "t_ns", "stopped"}``;
        FINAL/LOG carry the unpickled blob in ``data``."""
        if self.pump:
            n = self.lib.mgr_next(self.h, self._buf, 256, int(timeout_s * 1e6))
        else:
            n = self.lib.mgr_poll(self.h, self._buf, 256)
            if n == 0 and timeout_s > 0:
                time.sleep(min(timeout_s, 0.0005))
        out = []
        for i in range(n):
            r = self._buf[i]
            msg = {"type": REC_NAMES.get(r.type, "ERR"), "partition_id": int(r.worker),
                   "trial_id": f"{r.trial_id:016x}" if r.trial_id else None, "step": int(r.step),
                   "value": float(r.value), "t_ns": int(r.t_ns), "aux": int(r.aux),
                   "stopped": bool(r.reserved[0]), "data": None}
            if r.type in (REC_FINAL, REC_LOG, REC_ERROR) and r.aux > 0:
                k = self.lib.mgr_read_blob(self.h, r.worker, self._blob, len(self._blob))
                if k > 0:
                    raw = self._blob.raw[:k]
                    try:
                        msg["data"] = raw.decode("utf-8", "replace") if r.type == REC_LOG else cloudpickle.loads(raw)
                    except Exception as e:  # noqa: BLE001 - one undecodable blob must not lose the popped batch
                        msg["data"] = None
                        msg["decode_error"] = f"{type(e).__name__}: {e}"

            if msg["type"] == "REG":
                self.reservations.add({"partition_id": msg["partition_id"], "task_attempt": msg["aux"],
                                       "host_port": f"shm:{self.name}#{msg['partition_id']}",
                                       "pid": self.worker_pid(msg["partition_id"]), "trial_id": None})
            elif msg["type"] == "FINAL":
                self.reservations.assign_trial(msg["partition_id"], None)
            out.append(msg)
        return out

    # -- send ------------------------------------------------------------------------------------
    def assign(self, worker: int, trial_id: str, payload, tag: int = 0) -> bool:
        """Place a trial in the worker's assign slot. False if the previous one was not picked up yet. ``tag`` (ASHA:
        rung + 1) is echoed in the trial's FINAL record for the device scheduler (mg_ring.h ``assign_tag``)."""
        blob = cloudpickle.dumps(payload)
        if len(blob) > self.payload_bytes:
            raise ValueError(f"trial payload of {len(blob)} bytes exceeds the {self.payload_bytes}-byte assign slot")
        rc = self.lib.mgr_assign_tagged(self.h, worker, int(trial_id, 16), blob, len(blob), int(tag))
        if rc == -2:
            return False
        if rc != 0:
            raise OSError(f"mgr_assign failed: {rc}")
        self.reservations.assign_trial(worker, trial_id)
        return True

    def stage(self, worker: int, trial_id: str, payload, tag: int = 0) -> bool:
        """Park ``worker``'s NEXT trial in its staged slot (prefetch): the device scheduler — or :meth:`publish_staged`
        — hands it over the moment the worker's FINAL shows up. False if a trial is already parked there."""
        blob = cloudpickle.dumps(payload)
        if len(blob) > self.payload_bytes:
            raise ValueError(f"trial payload of {len(blob)} bytes exceeds the {self.payload_bytes}-byte staged slot")
        rc = self.lib.mgr_stage(self.h, worker, int(trial_id, 16), blob, len(blob), int(tag))
        if rc not in (0, -2):
            raise OSError(f"mgr_stage failed: {rc}")
        return rc == 0

    def staged_state(self, worker: int):
        """``("parked" | "published" | None, trial_id_hex)`` for the worker's staged slot."""
        tid = ctypes.c_uint64()
        rc = int(self.lib.mgr_staged_state(self.h, worker, ctypes.byref(tid)))
        return ({1: "parked", 2: "published"}.get(rc), f"{tid.value:016x}")

    def publish_staged(self, worker: int) -> bool:
        ok = int(self.lib.mgr_publish_staged(self.h, worker)) == 0
        if ok:
            self.reservations.assign_trial(worker, self.staged_state(worker)[1])
        return ok

    def device_decision(self, worker: int, expected_seq: int, timeout_s: float = 0.002):
        """The device scheduler's ``expected_seq``-th next-trial decision for ``worker`` (mg_asha.h) as
        ``(kind, rung, parent_trial_id_hex, t_ns)``, or None if K9 has not published it within ``timeout_s``."""
        kind, rung, parent, t_ns = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint64(), ctypes.c_uint64()
        deadline = time.perf_counter() + timeout_s
        while True:
            seq = int(self.lib.mgr_decision(self.h, worker, ctypes.byref(kind), ctypes.byref(rung), ctypes.byref(parent),
                                            ctypes.byref(t_ns)))
            if seq == expected_seq:
                return int(kind.value), int(rung.value), f"{parent.value:016x}", int(t_ns.value)
            if seq > expected_seq or time.perf_counter() > deadline:
                return None

    def stop_trial(self, worker: int, trial_id: str):
        self.lib.mgr_stop_trial(self.h, worker, int(trial_id, 16))

    def gstop(self):
        self.lib.mgr_gstop(self.h)

    def configure_policy(self, enabled: bool, direction: str, es_min: int, es_interval: int):
        """Arm the native median-stopping-rule engine (csrc/ctrl/ring.cpp policy_on_metric)."""
        self.lib.mgr_policy_config(self.h, 1 if enabled else 0, 1 if direction == "max" else -1, es_min,
                                   max(1, es_interval))

    # -- liveness --------------------------------------------------------------------------------
    def heartbeat_age_s(self, worker: int) -> float:
        hb = int(self.lib.mgr_worker_heartbeat(self.h, worker))
        return 0.0 if hb == 0 else max(0.0, (now_ns() - hb) / 1e9)  # 0: has not started beating yet

    def worker_state(self, worker: int) -> int:
        return int(self.lib.mgr_worker_state(self.h, worker))

    def worker_pid(self, worker: int) -> int:
        return int(self.lib.mgr_worker_pid(self.h, worker))

    def policy_stops(self) -> int:
        return int(self.lib.mgr_policy_stops(self.h))

    def stop_time_ns(self, worker: int) -> int:
        return int(self.lib.mgr_stop_time_ns(self.h, worker))

    # -- the reference Server's method names (rpc.py:282-305, 335-392, 504-511) ---------------------------------------
    def start(self, exp_driver=None):
        """The segment exists and its pump runs from construction on; returns the "address" workers attach to
        (reference: ``Server.start`` returns ``(host, port)``)."""
        return self.name, self.num_workers

    def stop(self):
        self.close()

    def await_reservations(self, sc=None, status=None, timeout=600):
        """Block until every worker has registered (messages are drained by whoever calls :meth:`next`)."""
        deadline = time.time() + timeout
        while not self.reservations.done():
            if time.time() > deadline:
                raise TimeoutError(f"{self.reservations.remaining()} workers did not register within {timeout} s")
            time.sleep(0.01)
        return self.reservations.get()

    def get_assigned_trial_id(self, partition_id: int):
        return self.reservations.get_assigned_trial(partition_id)

    def close(self, unlink: bool = True):
        if self.h:
            self.lib.mgr_close(self.h, 1 if unlink else 0)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WorkerChannel:
    """Worker end (one per worker process)."""

    GSTOP = object()

    def __init__(self, name: str, worker_id: int, attempt: int = 0):
        self.lib = _host()
        self.name, self.worker, self.attempt = name, worker_id, attempt
        self.h = None
        deadline = time.time() + 30
        while not self.h:
            self.h = self.lib.mgr_attach(name.encode())
            if not self.h:
                if time.time() > deadline:
                    raise OSError(f"could not attach control segment {name}")
                time.sleep(0.01)
        # the driver may place payloads up to the segment's slot size (pickled model / dataset closures of an
        # ablation trial): size the receive buffer from the header, not from a constant
        self._payload = ctypes.create_string_buffer(max(1 << 16, int(self.lib.mgr_payload_bytes(self.h))))
        self._dev = None  # lazily created device-side producer state
        self._current_trial = None  # last trial handed out by get_suggestion()
        self._hb = None

    def register(self):
        rc = self.lib.mgr_register(self.h, self.worker, self.attempt, os.getpid())
        if rc != 0:
            raise OSError(f"mgr_register failed: {rc}")

    def set_state(self, state: int):
        self.lib.mgr_set_state(self.h, self.worker, state)

    def heartbeat(self):
        self.lib.mgr_heartbeat(self.h, self.worker)

    def post_metric(self, trial_id: str, value: float, step: int):
        self.lib.mgr_post(self.h, self.worker, REC_METRIC, int(trial_id, 16), int(step), float(value), 0)

    def post_idle(self):
        self.lib.mgr_post(self.h, self.worker, REC_IDLE, 0, 0, 0.0, 0)

    def post_exit(self):
        self.lib.mgr_post(self.h, self.worker, REC_EXIT, 0, 0, 0.0, 0)

    def _post_blob(self, rec_type, trial_id, step, value, blob: bytes):
        rc = self.lib.mgr_post_blob(self.h, self.worker, rec_type, int(trial_id, 16) if trial_id else 0, int(step),
                                    float(value), blob, len(blob), 30_000_000)
        if rc != 0:
            raise OSError(f"mgr_post_blob failed: {rc}")

    def post_final(self, trial_id: str, metric, payload: Optional[dict] = None, step: int = 0):
        data = {"metric": metric}
        if payload:
            data.update(payload)
        v = float(metric) if isinstance(metric, (int, float)) else float("nan")
        self._post_blob(REC_FINAL, trial_id, step, v, cloudpickle.dumps(data))

    def post_error(self, trial_id: Optional[str], text: str):
        self._post_blob(REC_ERROR, trial_id, 0, float("nan"), cloudpickle.dumps({"error": text}))

    def post_log(self, trial_id: Optional[str], text: str):
        if text:
            self._post_blob(REC_LOG, trial_id, 0, 0.0, text.encode("utf-8")[-(1 << 19):])

    def wait_assign(self, timeout_s: float = 1.0):
        """-> (trial_id, payload) | None (timeout) | WorkerChannel.GSTOP."""
        tid, n = ctypes.c_uint64(0), ctypes.c_uint32(0)
        rc = self.lib.mgr_wait_assign(self.h, self.worker, int(timeout_s * 1e6), ctypes.byref(tid), self._payload,
                                      len(self._payload), ctypes.byref(n))
        if rc == 1:
            payload = cloudpickle.loads(self._payload.raw[:n.value]) if n.value else None
            return f"{tid.value:016x}", payload
        if rc == 2:
            return WorkerChannel.GSTOP
        if rc == 0:
            return None
        raise OSError(f"mgr_wait_assign failed: {rc}")

    def check_stop(self, trial_id: str) -> bool:
        return bool(self.lib.mgr_check_stop(self.h, self.worker, int(trial_id, 16)))

    def gstop_set(self) -> bool:
        return bool(self.lib.mgr_gstop_set(self.h))

    # -- device path (K8) ------------------------------------------------------------------------
    def _device_state(self):
        if self._dev is None:
            import torch

            cu = _lib.cuda()
            cu.mg_set_device(torch.cuda.current_device())
            u64, P = ctypes.c_ulonglong, ctypes.c_void_p
            cu.mg_ring_map.argtypes, cu.mg_ring_map.restype = [P, u64], u64
            cu.mg_report_store.argtypes = [u64, u64, u64, ctypes.c_uint, P, ctypes.c_int, ctypes.c_longlong, u64,
                                           ctypes.c_uint, P, u64, ctypes.c_uint, ctypes.c_uint, ctypes.c_longlong, P]
            cu.mg_report_store.restype = ctypes.c_int
            cu.mg_mailbox_open.argtypes, cu.mg_mailbox_open.restype = [ctypes.c_char_p], u64
            cu.mg_mailbox_close.argtypes, cu.mg_mailbox_close.restype = [u64], ctypes.c_int
            cu.mg_gt_offset_ns.argtypes, cu.mg_gt_offset_ns.restype = [], ctypes.c_longlong
            base = self.lib.mgr_base(self.h)
            dev_base = cu.mg_ring_map(base, self.lib.mgr_bytes(self.h))
            if not dev_base:
                raise RuntimeError("cudaHostRegister of the control segment failed")
            # device mailbox: the scheduler GPU's peer-visible ring (published by the driver when K9 runs); without
            # it, or when this GPU cannot map it, the host-mapped device ring alone carries the records
            mailbox, mbox_slots = 0, 0
            handle = ctypes.create_string_buffer(64)
            slots_out = ctypes.c_uint32(0)
            if os.environ.get("MAGGY_B200_MAILBOX", "1") != "0" and \
                    self.lib.mgr_get_mailbox(self.h, handle, ctypes.byref(slots_out)):
                mailbox = int(cu.mg_mailbox_open(handle.raw))
                mbox_slots = int(slots_out.value) if mailbox else 0
            self._dev = {
                "cu": cu, "base": dev_base,
                "off_ring": int(self.lib.mgr_off_dev_ring(self.h, self.worker)),
                "off_ctl": int(self.lib.mgr_off_ctl(self.h, self.worker)),
                "slots": int(self.lib.mgr_ring_slots(self.h)),
                "state": torch.zeros(8, dtype=torch.int32, device="cuda"),
                "mailbox": mailbox, "mbox_slots": mbox_slots, "mbox_workers": int(self.lib.mgr_num_workers(self.h)),
                "gt_offset": int(cu.mg_gt_offset_ns()),
            }
        return self._dev

    def post_metric_device(self, trial_id: str, tensor, step: int):
        """Publish a 0-dim CUDA tensor as a METRIC record from the GPU (no host synchronisation)."""
        import torch

        d = self._device_state()
        d["cu"].mg_set_device(torch.cuda.current_device())
        code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2, torch.float64: 3}.get(tensor.dtype)
        if code is None:
            tensor, code = tensor.float(), 0
        rc = d["cu"].mg_report_store(d["base"], d["off_ring"], d["off_ctl"], d["slots"], tensor.data_ptr(), code,
                                     int(step), int(trial_id, 16), self.worker, d["state"].data_ptr(),
                                     d["mailbox"], d["mbox_workers"], d["mbox_slots"], d["gt_offset"],
                                     torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"mg_report_store failed: {rc}")

    # -- the reference Client's method names (rpc.py:708-802) ------------------------------------------------------
    def await_reservations(self):
        """Nothing to wait for: rendezvous facts of gang / distributed runs travel in the assign payload."""
        return True

    def start_heartbeat(self, reporter, interval_s: float = 1.0):
        """Sign-of-life thread (the reference's heartbeat also carried metrics; here they are ring records)."""
        from .executors.trial_executor import _start_heartbeat

        self._hb = _start_heartbeat(self, interval_s, reporter)
        return self._hb

    def get_suggestion(self, reporter=None, timeout_s: float = None):
        """Block until the driver (or the device scheduler) hands over the next trial: ``(trial_id, params)``;
        ``(None, None)`` once the experiment is over (reference: polls GET with ``sleep(1)``, rpc.py:739-748)."""
        deadline = None if timeout_s is None else time.time() + timeout_s
        while deadline is None or time.time() < deadline:
            got = self.wait_assign(1.0)
            if got is WorkerChannel.GSTOP:
                return None, None
            if got is not None:
                trial_id, payload = got
                self._current_trial = trial_id
                return trial_id, (payload or {}).get("params", payload)
        return None, None

    def get_message(self, msg_type, timeout=60):
        """The request / reply pairs of the TCP protocol that still have an answer on the shared segment."""
        if msg_type == "GET":
            return self.get_suggestion(timeout_s=timeout)
        if msg_type == "QUERY":
            return True   # reservations complete: the pool only assigns trials to registered workers
        raise NotImplementedError(f"{msg_type}: no request/reply exchange on the shared-memory control plane")

    def finalize_metric(self, metric, reporter=None):
        """FINAL + reporter reset (reference: rpc.py:793-802 holds the reporter lock across both)."""
        if reporter is None:
            return self.post_final(self._current_trial, metric)
        with reporter.lock:
            _, _, logs = reporter.get_data()
            self.post_final(reporter.trial_id or self._current_trial, metric, {"logs": logs} if logs else None)
            reporter.reset()

    def stop(self):
        if self._hb is not None:
            self._hb.set()   # the heartbeat thread's stop event
        self.close()

    def close(self):
        if self.h:
            if self._dev is not None:
                try:
                    import torch

                    torch.cuda.synchronize()
                    cu = self._dev["cu"]
                    if self._dev.get("mailbox"):
                        cu.mg_mailbox_close(self._dev["mailbox"])
                    cu.mg_ring_unmap.argtypes, cu.mg_ring_unmap.restype = [ctypes.c_void_p], ctypes.c_int
                    cu.mg_ring_unmap(self.lib.mgr_base(self.h))
                except Exception:
                    pass
                self._dev = None
            self.lib.mgr_close(self.h, 0)
            self.h = None


# ------------------------------------------------------------------------------------------------
# names of the reference's maggy/core/rpc.py. The TCP server / client pair does not exist
# here — one shared-memory segment replaces the sockets — so the reference's class names resolve to the object that plays
# the same role on this control plane.
# ------------------------------------------------------------------------------------------------
from .reservations import MirroredReservations, Reservations  # noqa: E402,F401  (rpc.py:45-202 in the reference)

Server = DriverChannel                     # rpc.py:260-392: listener + dispatch -> DriverChannel.next() / message callbacks
OptimizationServer = DriverChannel         # rpc.py:395-511: REG / QUERY / METRIC / FINAL / GET / LOG
DistributedTrainingServer = DriverChannel  # rpc.py:514-590: + EXEC_CONFIG (travels in the assign payload of gang trials)
TensorflowServer = DriverChannel           # rpc.py:593-633: + TF_CONFIG / RESERVATIONS (MirroredReservations.get_spec)
Client = WorkerChannel                     # rpc.py:636-802: heartbeat + GET polling -> records + futex-waited assign slot


class MessageSocket:
    """rpc.py:205-257 framed cloudpickled dicts over TCP. The wire format here is the 64-byte :class:`MgRecord` plus the
    per-worker blob areas; ``send`` / ``receive`` map onto a channel's post / next."""

    Record = MgRecord

    @staticmethod
    def send(channel, msg: dict):
        kind = msg.get("type")
        if kind == "METRIC":
            return channel.post_metric(msg["trial_id"], msg["data"]["value"], msg["data"].get("step", 0))
        if kind == "FINAL":
            return channel.post_final(msg["trial_id"], msg["data"])
        if kind == "LOG":
            return channel.post_log(msg.get("trial_id"), msg.get("data", ""))
        raise NotImplementedError(f"message type {kind}")

    @staticmethod
    def receive(channel, timeout_s: float = 0.1):
        return channel.next(timeout_s)
