"""Host handle of the persistent device scheduler (K9, csrc/ctrl/ring_device.cu ``sched_kernel``).

``DeviceScheduler(channel, ...).start()`` maps the control segment into the CUDA address space and launches
a single-CTA kernel on a low-priority stream that keeps polling every worker's rings, evaluates the median
stopping rule on the device and writes STOP words itself. ``stop()`` raises the quit flag, joins the stream
and returns the kernel's counters. The Python driver keeps consuming the same records for bookkeeping.
"""
from __future__ import annotations

import atexit
import ctypes

import torch

from ..ops import _lib
from . import rpc


class DeviceScheduler:
    def __init__(self, channel: "rpc.DriverChannel", direction: str = "max", es_min: int = 10, es_interval: int = 1,
                 max_steps: int = 1024, cap_final: int = 1024, device="cuda:0", use_mailbox: bool = True,
                 asha: dict = None):
        """``asha={"rf": 2, "max_rung": 2, "cap": 256}``: K9 also keeps the ASHA rung tables and takes the next-trial
        decision (promote which configuration / start a fresh one) when it sees a worker's FINAL record
        (csrc/include/mg_asha.h; reference: maggy/optimizer/asha.py:95-147 on the driver)."""
        self.asha = asha
        self.ch = channel
        self.device = torch.device(device)
        self.direction = 1 if direction == "max" else -1
        self.es_min, self.es_interval = es_min, es_interval
        self.max_steps, self.cap_final = max_steps, cap_final
        self.cu = _lib.cuda()
        u64, P, u32, i32 = ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_uint, ctypes.c_int
        self._u64 = u64
        self.cu.mg_ring_map.argtypes, self.cu.mg_ring_map.restype = [P, u64], u64
        self.cu.mg_sched_scratch_bytes.argtypes, self.cu.mg_sched_scratch_bytes.restype = [u32, u32, u32], u64
        self.cu.mg_sched_start.argtypes = [u64, u64, u64, u64, u64, u32, u32, i32, u32, u32, u32, u32, P, u64, u64, u32,
                                           ctypes.c_longlong, P, P]
        self.cu.mg_asha_table_bytes.argtypes, self.cu.mg_asha_table_bytes.restype = [u32], u64
        self.asha_table = None
        self.cu.mg_sched_start.restype = i32
        self.cu.mg_mailbox_create.argtypes, self.cu.mg_mailbox_create.restype = [u32, u32, ctypes.c_char_p], u64
        self.cu.mg_mailbox_free.argtypes, self.cu.mg_mailbox_free.restype = [u64], i32
        self.cu.mg_gt_offset_ns.argtypes, self.cu.mg_gt_offset_ns.restype = [], ctypes.c_longlong
        self.use_mailbox = use_mailbox
        self.mailbox = 0
        self.mailbox_slots = 256
        self.stream = None
        self.running = False

    def start(self):
        host = self.ch.lib
        h = self.ch.h
        with torch.cuda.device(self.device):
            self._host_base = host.mgr_base(h)
            base = self.cu.mg_ring_map(self._host_base, host.mgr_bytes(h))
            if not base:
                raise RuntimeError("could not map the control segment into the CUDA address space")
            W = self.ch.num_workers
            nbytes = int(self.cu.mg_sched_scratch_bytes(W, self.max_steps, self.cap_final))
            self.scratch = torch.zeros(nbytes + 64, dtype=torch.uint8, device=self.device)
            # quit flag lives in pinned host memory the kernel polls through its device mapping
            self.quit = torch.zeros(16, dtype=torch.int32).pin_memory()
            quit_dev = self.quit.data_ptr()  # torch-pinned memory is UVA-addressable from the device as is
            lo, _hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, 0)
            self.stream = torch.cuda.Stream(device=self.device, priority=0)
            u64 = self._u64
            host.mgr_off_host_ring.argtypes, host.mgr_off_host_ring.restype = [ctypes.c_void_p, ctypes.c_uint32], u64
            host.mgr_ring_stride.argtypes, host.mgr_ring_stride.restype = [ctypes.c_void_p], u64
            off_ctl = int(host.mgr_off_ctl(h, 0))
            off_dev0 = int(host.mgr_off_dev_ring(h, 0))
            off_host0 = int(host.mgr_off_host_ring(h, 0))
            ring_stride = int(host.mgr_ring_stride(h))
            self.cu.mg_set_device(self.device.index or 0)
            if self.use_mailbox:
                # peer-visible record rings + STOP words in THIS GPU's memory; the handle travels through the segment
                handle = ctypes.create_string_buffer(64)
                self.mailbox = int(self.cu.mg_mailbox_create(W, self.mailbox_slots, handle))
                if self.mailbox:
                    host.mgr_set_mailbox(h, handle.raw, self.mailbox_slots,
                                         W * (64 + 64 * self.mailbox_slots))
            self.gt_offset = int(self.cu.mg_gt_offset_ns())
            asha_ptr = None
            if self.asha:
                import struct

                cap = int(self.asha.get("cap", 256))
                nbytes = int(self.cu.mg_asha_table_bytes(cap))
                host_tab = torch.zeros(nbytes, dtype=torch.uint8)
                hdr = struct.pack("<IIIi", int(self.asha["rf"]), int(self.asha["max_rung"]), cap, self.direction)
                host_tab[:len(hdr)] = torch.frombuffer(bytearray(hdr), dtype=torch.uint8)
                self.asha_table = host_tab.to(self.device)
                torch.cuda.synchronize(self.device)
                asha_ptr = self.asha_table.data_ptr()
            rc = self.cu.mg_sched_start(base, off_ctl, off_host0, off_dev0, ring_stride, W,
                                        int(host.mgr_ring_slots(h)), self.direction, self.es_min, self.es_interval,
                                        self.max_steps, self.cap_final, self.scratch.data_ptr(), quit_dev,
                                        self.mailbox, self.mailbox_slots if self.mailbox else 0, self.gt_offset,
                                        asha_ptr, self.stream.cuda_stream)
            if rc != 0:
                raise RuntimeError(f"mg_sched_start failed: {rc}")
        self.running = True
        # a persistent kernel that outlives its owner keeps the CUDA context (and the process) from shutting down
        atexit.register(self._raise_quit)
        return self

    def _raise_quit(self):
        try:
            self.quit[0] = 1
        except Exception:  # noqa: BLE001
            pass

    def __del__(self):
        self._raise_quit()

    def stats(self) -> dict:
        W = self.ch.num_workers
        off = 4 * self.cap_final * self.max_steps + 4 * self.cap_final + 8 + 4 * W * self.max_steps
        off = (off + 7) // 8 * 8 + 8 * W * 3
        raw = self.scratch[off:off + 56].cpu().view(torch.int64).tolist()
        return {"polls": raw[0], "metrics": raw[1], "stops": raw[2], "finals": raw[3], "mailbox_records": raw[4],
                "asha_decisions": raw[5], "handoffs": raw[6], "mailbox": bool(self.mailbox)}

    def stop(self) -> dict:
        if self.running:
            self.quit[0] = 1
            self.stream.synchronize()
            self.running = False
            stats = self.stats()
            self.cu.mg_ring_unmap.argtypes, self.cu.mg_ring_unmap.restype = [ctypes.c_void_p], ctypes.c_int
            self.cu.mg_ring_unmap(self._host_base)  # never leave a registration behind an unmapped segment
            if self.mailbox:
                # workers that still hold the IPC mapping keep the allocation alive until they close it
                self.ch.lib.mgr_clear_mailbox(self.ch.h)
                self.cu.mg_mailbox_free(self.mailbox)
                self.mailbox = 0
            return stats
        return self.stats()
