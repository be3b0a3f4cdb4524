"""ZeRO-3 style sharded weights whose all-gather is fused into the consuming GEMM (K3, csrc/comm/ag_gemm.cu).

Reference: fairscale FSDP all-gathers a wrapped module's flat parameter with NCCL, then runs the cuBLAS GEMM
(maggy/core/patching/modules.py:75-95). ``ShardedWeight`` keeps only ``rows / world`` rows of a ``[rows, K]``
weight per rank in symmetric memory; ``ag_linear(x, w)`` runs ONE kernel in which a few gather CTAs stream
the remote shards over NVLink while the tcgen05 pipeline already multiplies against the local shard.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from ..ops import _lib
from ..ops import functional as F

_registered = False


def _register():
    global _registered
    if not _registered:
        P, u64p, i, ll, u = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_uint
        _lib.register_cuda_symbol("mg_ag_gemm", [P, P, P, u64p, P, P, u64p, P, i, i, i, ll, ll, i, i, i, u, P])
        _registered = True


class ShardedWeight:
    """Rows ``[rank*rows/world, (rank+1)*rows/world)`` of a ``[rows, K]`` bf16 weight, in symmetric memory."""

    def __init__(self, full_weight: torch.Tensor, group=None, device=None, n_copy: int = 0):
        import torch.distributed._symmetric_memory as symm_mem

        _register()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        rows, K = full_weight.shape
        assert rows % self.world == 0 and (rows // self.world) % 256 == 0, "rows per shard must be a multiple of 256"
        assert K % 8 == 0
        self.rows, self.K = rows, K
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        gname = group.group_name if group is not None else dist.group.WORLD.group_name
        per = rows // self.world
        self.shard = symm_mem.empty((per, K), dtype=torch.bfloat16, device=self.device)
        self.shard.copy_(full_weight[self.rank * per:(self.rank + 1) * per].to(self.device, torch.bfloat16))
        self._h_shard = symm_mem.rendezvous(self.shard, gname)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=self.device)
        self.flags.zero_()
        self._h_flags = symm_mem.rendezvous(self.flags, gname)
        self.gathered = torch.empty((rows, K), dtype=torch.bfloat16, device=self.device)
        self.ready = torch.zeros(rows // 256, dtype=torch.int32, device=self.device)
        self.local_sync = torch.zeros(4, dtype=torch.int32, device=self.device)
        arr = ctypes.c_ulonglong * self.world
        self._shard_ptrs = arr(*[int(p) for p in self._h_shard.buffer_ptrs])
        self._flag_ptrs = arr(*[int(p) for p in self._h_flags.buffer_ptrs])
        self.n_copy = n_copy
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)


def ag_linear(x: torch.Tensor, w: ShardedWeight, bias: Optional[torch.Tensor] = None, out=None) -> torch.Tensor:
    """``x[..., K] @ W_fullᵀ (+ bias)`` with the all-gather of W fused into the GEMM kernel."""
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, K = x2.shape
    assert K == w.K and x2.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((M, w.rows), dtype=torch.bfloat16, device=x.device)
    w.epoch += 1
    rc = _lib.cuda().mg_ag_gemm(x2.data_ptr(), out.data_ptr(), 0 if bias is None else bias.data_ptr(), w._shard_ptrs,
                                w.gathered.data_ptr(), w.ready.data_ptr(), w._flag_ptrs, w.local_sync.data_ptr(),
                                M, w.rows, K, x2.stride(0), out.stride(0), w.rank, w.world, w.n_copy, w.epoch,
                                F._stream())
    _lib.check(rc, "mg_ag_gemm")
    F._count(1)
    return out.view(*x.shape[:-1], w.rows)
