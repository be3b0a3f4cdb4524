"""Per-GPU worker process pool (replaces the Spark executor layer: reference
maggy/core/experiment_driver/spark_driver.py:136-145 ``foreachPartition(executor_fn)``).

Each *slot* is one trial executor: a gang of ``gpus_per_worker`` processes (1 by default) pinned to its
own GPUs with ``CUDA_VISIBLE_DEVICES`` and to a CPU-affinity range, started by the native launcher
(csrc/runtime/launcher.cpp, posix_spawn, own process group). Gang rank 0 ("leader") owns the slot's
control ring; ranks 1.. follow the leader through a per-gang torch.distributed group. The driver
supervises: unexpected exits are detected by ``poll()`` and the slot is respawned with ``attempt + 1``
which makes the driver reschedule the lost trial (reference semantics: rpc.py:421-432).
"""
from __future__ import annotations

import ctypes
import os
import signal
import socket
import sys
from typing import Dict, List, Optional

from ..ops import _lib

_cfg = False


def _launcher():
    global _cfg
    lib = _lib.host()
    if not _cfg:
        P, i = ctypes.c_void_p, ctypes.c_int
        lib.mgl_create.argtypes, lib.mgl_create.restype = [i], P
        lib.mgl_destroy.argtypes, lib.mgl_destroy.restype = [P], None
        lib.mgl_spawn.argtypes = [P, i, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, i, i, ctypes.c_char_p]
        lib.mgl_spawn.restype = i
        for name in ("mgl_poll", "mgl_pid", "mgl_exit_code", "mgl_term_signal", "mgl_spawn_count"):
            getattr(lib, name).argtypes, getattr(lib, name).restype = [P, i], i
        lib.mgl_wait.argtypes, lib.mgl_wait.restype = [P, i, i], i
        lib.mgl_kill.argtypes, lib.mgl_kill.restype = [P, i, i], i
        _cfg = True
    return lib


def _blob(items: List[str]) -> bytes:
    return b"".join(s.encode() + b"\0" for s in items) + b"\0"


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class WorkerPool:
    def __init__(self, num_workers: int, segment: str, spec_path: str, log_dir: str,
                 gpus_per_worker: int = 1, use_gpus: Optional[bool] = None, extra_env: Optional[Dict[str, str]] = None,
                 share_devices: bool = False, also_visible: Optional[List[int]] = None):
        from .. import util

        self.lib = _launcher()
        self.num_workers = num_workers
        self.gang = max(1, gpus_per_worker)
        self.segment, self.spec_path, self.log_dir = segment, spec_path, log_dir
        self.pool = self.lib.mgl_create(num_workers * self.gang)
        ngpu = util.num_physical_devices()
        self.use_gpus = (ngpu >= num_workers * self.gang) if use_gpus is None else use_gpus
        self.attempts = [0] * num_workers
        self.ports = [free_port() for _ in range(num_workers)]
        self.extra_env = dict(extra_env or {})
        self.ncpu = os.cpu_count() or 1
        # share_devices: every worker sees ALL pool GPUs and selects its own by MAGGY_LOCAL_RANK (needed when
        # the workers form one collective job: peer-mapped symmetric memory wants distinct device indices)
        self.share_devices = share_devices
        # GPUs every worker must additionally SEE (appended after its own, so cuda:0 stays its training device): the
        # device-scheduler GPU, whose mailbox a worker can only map through CUDA IPC if that device is visible to it
        self.also_visible = [int(g) for g in (also_visible or [])]

    def _slot(self, worker: int, rank: int) -> int:
        return worker * self.gang + rank

    def spawn(self, worker: int):
        """(Re)start every process of worker ``worker``."""
        attempt = self.attempts[worker]
        cpus_per = max(1, self.ncpu // max(1, self.num_workers * self.gang))
        for rank in range(self.gang):
            slot = self._slot(worker, rank)
            argv = [sys.executable, "-m", "maggy_b200.core.executors.worker_main",
                    "--segment", self.segment, "--worker-id", str(worker), "--attempt", str(attempt),
                    "--spec", self.spec_path, "--gang-rank", str(rank), "--gang-size", str(self.gang),
                    "--gang-port", str(self.ports[worker])]
            # the worker must be able to import whatever module train_fn was defined in (cloudpickle pickles
            # importable functions by reference), so it inherits the driver's sys.path
            repo_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            paths = [repo_root] + [p for p in sys.path if p] + \
                    [p for p in os.environ.get("PYTHONPATH", "").split(os.pathsep) if p]
            env = {"PYTHONPATH": os.pathsep.join(dict.fromkeys(paths)), "MAGGY_WORKER_ID": str(worker),
                   "PYTHONUNBUFFERED": "1"}
            env.update(self.extra_env)
            if not self.use_gpus:
                gpus, local_rank = "-1", 0
            elif self.share_devices:
                gpus, local_rank = ",".join(str(i) for i in range(self.num_workers * self.gang)), slot
            elif self.gang > 1:
                # a gang sees exactly its own GPUs; rank r of the gang uses cuda:r
                gpus, local_rank = ",".join(str(worker * self.gang + r) for r in range(self.gang)), rank
            else:
                gpus, local_rank = str(slot), 0
            if self.use_gpus and not self.share_devices and self.also_visible:
                mine = gpus.split(",")
                gpus = ",".join(mine + [str(g) for g in self.also_visible if str(g) not in mine])
            env["MAGGY_LOCAL_RANK"] = str(local_rank)
            log = os.path.join(self.log_dir, f"worker_{worker}_{rank}.stdout.log")
            rc = self.lib.mgl_spawn(self.pool, slot, _blob(argv), _blob([f"{k}={v}" for k, v in env.items()]),
                                    gpus.encode(), (slot * cpus_per) % self.ncpu, cpus_per, log.encode())
            if rc < 0:
                raise OSError(f"could not spawn worker {worker} rank {rank}: {rc}")

    def start(self):
        for w in range(self.num_workers):
            self.spawn(w)

    def alive(self, worker: int) -> bool:
        return all(self.lib.mgl_poll(self.pool, self._slot(worker, r)) == 1 for r in range(self.gang))

    def exit_info(self, worker: int):
        """[(exit_code, signal)] of the worker's processes that have exited."""
        out = []
        for r in range(self.gang):
            s = self._slot(worker, r)
            if self.lib.mgl_poll(self.pool, s) == 0:
                out.append((self.lib.mgl_exit_code(self.pool, s), self.lib.mgl_term_signal(self.pool, s)))
        return out

    def pid(self, worker: int, rank: int = 0) -> int:
        return self.lib.mgl_pid(self.pool, self._slot(worker, rank))

    def kill(self, worker: int, sig: int = signal.SIGKILL):
        for r in range(self.gang):
            self.lib.mgl_kill(self.pool, self._slot(worker, r), sig)

    def respawn(self, worker: int):
        """Kill whatever is left of the worker's gang and start it again with attempt + 1."""
        self.kill(worker)
        for r in range(self.gang):
            self.lib.mgl_wait(self.pool, self._slot(worker, r), 5000)
        self.attempts[worker] += 1
        self.ports[worker] = free_port()
        self.spawn(worker)

    def wait_all(self, timeout_s: float = 10.0) -> bool:
        ok = True
        for w in range(self.num_workers):
            for r in range(self.gang):
                if self.lib.mgl_wait(self.pool, self._slot(w, r), int(timeout_s * 1000)) == 1:
                    ok = False
        return ok

    def shutdown(self, grace_s: float = 5.0):
        if not self.pool:
            return
        if not self.wait_all(grace_s):
            for w in range(self.num_workers):
                self.kill(w, signal.SIGTERM)
            if not self.wait_all(2.0):
                for w in range(self.num_workers):
                    self.kill(w, signal.SIGKILL)
                self.wait_all(2.0)
        self.lib.mgl_destroy(self.pool)
        self.pool = None
