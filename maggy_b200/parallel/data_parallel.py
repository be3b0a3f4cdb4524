"""Data-parallel module wrapper on the fused NVLink path (replaces torch DDP + ZeRO for modules built
on maggy_b200 ops; reference: maggy/core/patching/modules.py:38-65).

``DataParallel(module)`` re-homes the parameters into symmetric-memory arenas (``FlatParams``); the
gradient all-reduce DDP would do in bucket hooks is instead *fused into the optimizer step*:
``FusedAdamW.step()`` = reduce-scatter + Adam + all-gather in one kernel (csrc/comm/rs_adam.cu).
Rank 0's initial parameters are the ones every rank starts from (DDP's constructor broadcast,
reference SURVEY 2.4 C3): all ranks seed identically, and the arena copy is verified with a checksum.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .arena import FlatParams, FusedAdamW


def arena_dtype_for(module: torch.nn.Module):
    """bf16 / fp32 when every floating-point parameter of ``module`` has that dtype, else None (no single arena)."""
    dts = {p.dtype for p in module.parameters()}
    if dts == {torch.bfloat16}:
        return torch.bfloat16
    if dts == {torch.float32}:
        return torch.float32
    return None


def peer_broadcast(flat: FlatParams, key: str = "params", root: int = 0):
    """K2: broadcast the ``root`` rank's arena to every rank with the peer-store / multimem.st kernel
    (csrc/comm/rs_sum.cu ``mg_peer_bcast``) — the constructor broadcast DDP does with NCCL."""
    import ctypes

    from ..ops import _lib
    from ..ops import functional as F

    P, u64 = ctypes.c_void_p, ctypes.c_ulonglong
    _lib.register_cuda_symbol("mg_peer_bcast", [P, P, u64, P, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_uint, P])
    W = flat.world_size
    if not hasattr(flat, "_bcast_state"):
        flags = flat._alloc("bcast_flags", 64, torch.int32)
        flags.zero_()
        local = torch.zeros(4, dtype=torch.int32, device=flat.device)
        torch.cuda.synchronize(flat.device)
        dist.barrier(group=flat.group)
        flat._bcast_state = {"flags": flags, "local": local, "epoch": 0,
                             "flag_ptrs": (u64 * W)(*flat.peer_ptrs("bcast_flags"))}
    st = flat._bcast_state
    st["epoch"] += 1
    buf = getattr(flat, key)
    rc = _lib.cuda().mg_peer_bcast((u64 * W)(*flat.peer_ptrs(key)), st["flag_ptrs"], flat.multicast_ptr(key) if W >= 4 else 0,
                                   st["local"].data_ptr(), buf.numel() * buf.element_size(), flat.rank, W, root,
                                   st["epoch"], F._stream())
    _lib.check(rc, "mg_peer_bcast")
    F._count(1)


class DataParallel(torch.nn.Module):
    """``module`` may be ANY torch module whose parameters share one dtype (bf16 or fp32): modules built on maggy_b200
    ops write their gradients straight into the arena (``main_grad``); for everything else torch autograd's ``.grad``
    tensors are adopted into the arena on the first step and accumulate in place afterwards
    (``FlatParams.adopt_autograd_grads``), so the one fused reduce-scatter + Adam + all-gather kernel serves both."""

    def __init__(self, module: torch.nn.Module, device=None, group=None, broadcast_params: bool = True, dtype=None):
        super().__init__()
        self.module = module
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        dtype = dtype or arena_dtype_for(module) or torch.bfloat16
        self.flat = FlatParams(module, device=device, world_size=world, rank=rank, group=group, dtype=dtype)
        if world > 1 and broadcast_params:
            # rank 0 is authoritative (DDP's constructor broadcast): one kernel of peer / multicast stores (K2)
            peer_broadcast(self.flat, "params", root=0)
            if dtype == torch.bfloat16:   # rebuild the fp32 master shard from the agreed bf16 values
                lo = rank * self.flat.shard_numel
                self.flat.master.copy_(self.flat.params[lo:lo + self.flat.shard_numel].float())
            for b in module.buffers():    # running statistics etc.: small, rare — a library broadcast is fine
                dist.broadcast(b, src=0, group=group)
        self._optimizer: Optional[FusedAdamW] = None

    def allreduce_grads(self):
        """For optimizers other than Adam(W) (which are fused with the reduction): average the whole gradient arena
        with one collective before the optimizer's step."""
        self.flat.adopt_autograd_grads()
        self.flat.zero_unwritten_grads()
        if self.flat.world_size > 1:
            dist.all_reduce(self.flat.grads, op=dist.ReduceOp.AVG, group=self.flat.group)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def make_optimizer(self, **kw) -> FusedAdamW:
        """The fused sharded AdamW bound to this module's arenas."""
        self._optimizer = FusedAdamW(self.flat, **kw)
        return self._optimizer

    def parameters(self, recurse: bool = True):
        return self.module.parameters(recurse)

    def named_parameters(self, *a, **k):
        return self.module.named_parameters(*a, **k)

    def state_dict(self, *a, **k):
        return self.module.state_dict(*a, **k)
