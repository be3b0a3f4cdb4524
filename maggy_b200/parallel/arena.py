"""Flat parameter / gradient arenas and the fused sharded AdamW that consumes them.

Design (B200-first, replaces DDP's Reducer buckets + ``ZeroRedundancyOptimizer``; reference:
maggy/core/patching/modules.py:44-63, maggy/core/patching/optim.py:28-45):

* every parameter of the module is a *view* into ONE bf16 parameter arena, laid out
  ``[weight-decay params | no-decay params | padding]``; every parameter also gets a ``main_grad`` view
  into ONE bf16 gradient arena that the wgrad kernels write directly (see ops/functional.py);
* with ``world_size > 1`` both arenas live in symmetric (peer-mapped, NVSwitch-multicast-bound) memory
  and rank ``r`` owns the contiguous shard ``[r*shard, (r+1)*shard)``: fp32 master/m/v exist only for
  that shard (ZeRO-1/2 footprint);
* ``FusedAdamW.step()`` is ONE kernel launch: gradient reduce-scatter (in-switch or peer loads) +
  scale + Adam + bf16 parameter all-gather (``csrc/comm/rs_adam.cu``), or the plain flat Adam kernel
  when ``world_size == 1`` (``csrc/ops/optim.cu``).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional

import torch

from ..ops import _lib
from ..ops import functional as F

_ALIGN = 128  # elements; keeps every parameter 256-byte aligned (TMA / 16-byte vector requirements)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def default_no_decay(name: str, p: torch.Tensor) -> bool:
    """GPT-2 style: biases / norms / 1-D tensors are not weight-decayed."""
    return p.dim() < 2


class FlatParams:
    """Re-homes a module's parameters into flat bf16 arenas (optionally symmetric memory)."""

    def __init__(self, module: torch.nn.Module, device=None, world_size: int = 1, rank: int = 0,
                 group=None, no_decay_fn=default_no_decay, use_symmetric: Optional[bool] = None, include=None,
                 dtype=torch.bfloat16):
        """``include(name, param) -> bool`` restricts the arena to a subset of the module's parameters (e.g. the
        replicated ones of a partly sharded model; the shards get a second, rank-local arena).

        ``dtype``: ``torch.bfloat16`` (default: bf16 compute arena + fp32 master shard) or ``torch.float32`` — generic
        torch modules keep fp32 parameters and receive fp32 autograd gradients; their arena IS the master copy."""
        assert dtype in (torch.bfloat16, torch.float32)
        self.dtype = dtype
        params = []
        seen = set()
        for name, p in module.named_parameters():
            if id(p) in seen or (include is not None and not include(name, p)):
                continue
            seen.add(id(p))
            params.append((name, p))
        decay = [(n, p) for n, p in params if not no_decay_fn(n, p)]
        nodecay = [(n, p) for n, p in params if no_decay_fn(n, p)]
        self.entries: List[tuple] = []  # (name, param, offset, numel)
        off = 0
        for n, p in decay:
            self.entries.append((n, p, off, p.numel()))
            off = _round_up(off + p.numel(), _ALIGN)
        self.decay_boundary = off
        for n, p in nodecay:
            self.entries.append((n, p, off, p.numel()))
            off = _round_up(off + p.numel(), _ALIGN)
        self.world_size, self.rank, self.group = world_size, rank, group
        self.numel = _round_up(max(off, _ALIGN), _ALIGN * world_size)
        self.shard_numel = self.numel // world_size
        self.device = torch.device(device) if device is not None else params[0][1].device
        self.symmetric = (world_size > 1) if use_symmetric is None else use_symmetric
        self.handles = {}

        # fp32 snapshot of the initial values (the source of the fp32 master shard)
        init = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        for n, p, o, k in self.entries:
            init[o:o + k].copy_(p.detach().reshape(-1).to(device=self.device, dtype=torch.float32))

        self.params = self._alloc("params", self.numel, dtype)
        self.grads = self._alloc("grads", self.numel, dtype)
        self.params.copy_(init.to(dtype))
        self.grads.zero_()
        lo = self.rank * self.shard_numel
        # bf16 arena: separate fp32 master shard. fp32 arena: the rank's slice of the arena itself
        self.master = init[lo:lo + self.shard_numel].clone() if dtype == torch.bfloat16 else \
            self.params[lo:lo + self.shard_numel]
        del init
        for n, p, o, k in self.entries:
            p.data = self.params[o:o + k].view(p.shape)
            p.main_grad = self.grads[o:o + k].view(p.shape)
            p._mg_written = False
            p._maggy_flat = self  # lets the scoped optimizer patch find the arena behind a parameter
            p.grad = None

    def _alloc(self, key: str, numel: int, dtype) -> torch.Tensor:
        if not self.symmetric:
            return torch.empty(numel, dtype=dtype, device=self.device)
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(numel, dtype=dtype, device=self.device)
        gname = self.group.group_name if self.group is not None else torch.distributed.group.WORLD.group_name
        self.handles[key] = symm_mem.rendezvous(t, gname)
        return t

    def parameters(self) -> Iterable[torch.nn.Parameter]:
        return (p for _, p, _, _ in self.entries)

    def start_step(self):
        """Mark every main_grad as unwritten: the first wgrad of the step overwrites (no memset).

        Parameters whose gradient comes from torch autograd (see :meth:`adopt_autograd_grads`) accumulate in
        place into their arena slice, so those slices are zeroed here instead."""
        for _, p, _, _ in self.entries:
            p._mg_written = bool(getattr(p, "_mg_autograd", False))
        for lo, hi in getattr(self, "_autograd_ranges", ()):
            self.grads[lo:hi].zero_()

    def adopt_autograd_grads(self):
        """Parameters that are not touched by the maggy_b200 wgrad kernels (BatchNorm, ATen linear, ...) receive a
        ``.grad`` from torch autograd. The first time that happens the value is moved into the arena and ``p.grad``
        is re-pointed AT the arena slice: from then on autograd's AccumulateGrad adds in place, so the fused
        optimizer sees every gradient in the one flat buffer without per-step copies."""
        changed = False
        for _, p, _, _ in self.entries:
            g = p.grad
            if g is None or g.data_ptr() == p.main_grad.data_ptr():
                continue
            if p._mg_written:
                p.main_grad.add_(g.to(p.main_grad.dtype))
            else:
                p.main_grad.copy_(g)
            p._mg_written = True
            p._mg_autograd = True
            p.grad = p.main_grad
            changed = True
        if changed:
            # contiguous arena ranges (alignment padding included) that start_step() has to zero
            ranges = []
            for i, (_, p, o, k) in enumerate(self.entries):
                if not getattr(p, "_mg_autograd", False):
                    continue
                hi = self.entries[i + 1][2] if i + 1 < len(self.entries) else _round_up(o + k, _ALIGN)
                if ranges and ranges[-1][1] == o:
                    ranges[-1][1] = hi
                else:
                    ranges.append([o, hi])
            self._autograd_ranges = [tuple(r) for r in ranges]

    def zero_unwritten_grads(self):
        """Zero the gradient of parameters that received no wgrad this step (unused parameters)."""
        for _, p, _, _ in self.entries:
            if not p._mg_written:
                p.main_grad.zero_()

    def peer_ptrs(self, key: str) -> List[int]:
        if not self.symmetric:
            return [getattr(self, key).data_ptr()]
        return [int(x) for x in self.handles[key].buffer_ptrs]

    def multicast_ptr(self, key: str) -> int:
        if not self.symmetric:
            return 0
        try:
            return int(self.handles[key].multicast_ptr or 0)
        except Exception:
            return 0


class FusedAdamW:
    """AdamW over a :class:`FlatParams` arena — one kernel per step, sharded across ranks.

    ``grad_scale`` multiplies the (summed) gradient before the update; for data parallelism pass
    ``1 / world_size`` (the default) to average. ``max_grad_norm`` enables global-norm clipping.
    """

    def __init__(self, flat: FlatParams, lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1,
                 grad_scale: Optional[float] = None, max_grad_norm: float = 0.0, use_multicast="auto",
                 num_ctas: int = 0):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.grad_scale = (1.0 / flat.world_size) if grad_scale is None else grad_scale
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        self.epoch = 0
        self.num_ctas = num_ctas
        dev = flat.device
        self.m = torch.zeros(flat.shard_numel, dtype=torch.float32, device=dev)
        self.v = torch.zeros(flat.shard_numel, dtype=torch.float32, device=dev)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=dev) if max_grad_norm > 0 else None
        # device-resident step / lr so a captured CUDA graph replays with a moving schedule
        self.graph_safe = False
        self.step_tensor = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr_tensor = torch.full((1,), lr, dtype=torch.float32, device=dev)
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        if flat.world_size > 1:
            assert flat.symmetric, "multi-rank FusedAdamW needs symmetric-memory arenas"
            self.flags = flat._alloc("flags", 64, torch.int32)
            self.flags.zero_()
            self.local_sync = torch.zeros(4, dtype=torch.int32, device=dev)
            W = flat.world_size
            arr = ctypes.c_ulonglong * W
            self._grad_ptrs = arr(*flat.peer_ptrs("grads"))
            self._param_ptrs = arr(*flat.peer_ptrs("params"))
            self._flag_ptrs = arr(*flat.peer_ptrs("flags"))
            if use_multicast == "auto":
                # measured (profiles/k1_roofline_*.json): at 2 ranks direct peer loads/stores beat the
                # NVSwitch multicast path (0.43 vs 0.65 ms on the GPT-2 small arena); in-switch reduction pays
                # off once a rank would otherwise pull from many peers
                import os

                env = os.environ.get("MAGGY_B200_MULTICAST")
                use_multicast = (env == "1") if env in ("0", "1") else W >= 4
            self._grad_mc = flat.multicast_ptr("grads") if use_multicast else 0
            self._param_mc = flat.multicast_ptr("params") if use_multicast else 0
            sig = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_ulonglong,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                   ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_longlong, ctypes.c_float, ctypes.c_float,
                   ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                   ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
            for sym in ("mg_rs_adam", "mg_rs_adam_f32"):
                _lib.register_cuda_symbol(sym, sig)
            # make sure every rank zeroed its flags before anyone signals
            torch.cuda.synchronize(dev)
            torch.distributed.barrier(group=flat.group)

    @property
    def uses_multicast(self) -> bool:
        return self.flat.world_size > 1 and bool(self._grad_mc) and bool(self._param_mc)

    def zero_grad(self, set_to_none: bool = True):
        self.flat.start_step()

    def step(self, lr: Optional[float] = None):
        flat = self.flat
        lr = self.param_groups[0]["lr"] if lr is None else lr
        self.step_count += 1
        flat.adopt_autograd_grads()
        flat.zero_unwritten_grads()
        gn = None
        if self.max_grad_norm > 0:
            # local sum of squares of the local gradient; exact global norm needs the reduced gradient,
            # so for world > 1 we clip on the all-reduced norm of the *local* shards' contributions.
            if flat.dtype == torch.float32:
                self.gnorm_sq.zero_()
                _lib.check(_lib.cuda().mg_sumsq_f32(flat.grads.data_ptr(), flat.grads.numel(), self.gnorm_sq.data_ptr(),
                                                    F._stream()), "mg_sumsq_f32")
            else:
                F.sumsq(flat.grads, self.gnorm_sq)
            if flat.world_size > 1:
                torch.distributed.all_reduce(self.gnorm_sq, group=flat.group)
            gn = self.gnorm_sq
        b1, b2 = self.betas
        st = lt = None
        if self.graph_safe:
            self.step_tensor += 1  # captured as a tiny kernel: the replayed graph advances the step
            st, lt = self.step_tensor, self.lr_tensor
        if flat.world_size == 1 and flat.dtype == torch.float32:
            rc = _lib.cuda().mg_adam_flat_f32(flat.params.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                              flat.grads.data_ptr(), flat.numel, flat.decay_boundary, lr, b1, b2, self.eps,
                                              self.weight_decay, self.step_count, self.grad_scale,
                                              0 if gn is None else gn.data_ptr(), self.max_grad_norm,
                                              0 if st is None else st.data_ptr(), 0 if lt is None else lt.data_ptr(),
                                              F._stream())
            _lib.check(rc, "mg_adam_flat_f32")
            F._count(1)
        elif flat.world_size == 1:
            F.adam_flat_step(flat.master, self.m, self.v, flat.grads, flat.params,
                             decay_boundary=flat.decay_boundary, lr=lr, beta1=b1, beta2=b2, eps=self.eps,
                             weight_decay=self.weight_decay, step=self.step_count,
                             grad_scale=self.grad_scale, gnorm_sq=gn, max_norm=self.max_grad_norm,
                             step_tensor=st, lr_tensor=lt)
        else:
            self.epoch += 1
            kernel = _lib.cuda().mg_rs_adam_f32 if flat.dtype == torch.float32 else _lib.cuda().mg_rs_adam
            rc = kernel(
                self._grad_ptrs, self._param_ptrs, self._flag_ptrs, self._grad_mc, self._param_mc,
                flat.master.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.local_sync.data_ptr(),
                flat.shard_numel, flat.rank, flat.world_size, 0 if self.graph_safe else self.epoch,
                flat.decay_boundary, lr, b1, b2,
                self.eps, self.weight_decay, self.step_count, self.grad_scale,
                0 if gn is None else gn.data_ptr(), self.max_grad_norm,
                0 if st is None else st.data_ptr(), 0 if lt is None else lt.data_ptr(), self.num_ctas,
                F._stream())
            _lib.check(rc, "mg_rs_adam")
            F._count(1)
        flat.start_step()

    def current_step(self) -> int:
        """Number of optimizer steps taken. Under CUDA-graph replay only the device counter advances (the host
        ``step_count`` freezes at its capture-time value), so the device tensor is the source of truth there."""
        if self.graph_safe and getattr(self, "step_tensor", None) is not None:
            return int(self.step_tensor.item())
        return int(self.step_count)

    def state_dict(self):
        return {"step": self.current_step(), "m": self.m, "v": self.v, "master": self.flat.master,
                "rank": self.flat.rank, "world_size": self.flat.world_size}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        if getattr(self, "step_tensor", None) is not None:
            self.step_tensor.fill_(int(sd["step"]))
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.flat.master.copy_(sd["master"])
