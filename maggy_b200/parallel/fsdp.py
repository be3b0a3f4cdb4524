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


# ------------------------------------------------------------------------------------------------
# autograd integration: a linear layer whose weight lives sharded (ZeRO-3) and is gathered INSIDE the forward GEMM
# ------------------------------------------------------------------------------------------------
class _ShardedLinearFn(torch.autograd.Function):
    """``y = x W^T (+ b)`` with ``W`` row-sharded over the group.

    forward  : K3 ``ag_gemm`` — one kernel gathers the peer shards over NVLink while multiplying (the gathered copy
               stays in ``w.gathered`` for the backward of the same step);
    backward : dgrad on the gathered copy (tcgen05 GEMM, MN-major B), full wgrad GEMM, then a reduce-scatter so that
               each rank keeps only the gradient of its own rows (averaged over the group).
    """

    @staticmethod
    def forward(ctx, x, shard_param, bias, w: ShardedWeight):
        ctx.w = w
        ctx.has_bias = bias is not None
        ctx.bias_ref = bias
        ctx.shard_ref = shard_param
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ctx.save_for_backward(x2)
        ctx.xshape = x.shape
        return ag_linear(x2, w, bias).view(*x.shape[:-1], w.rows)

    @staticmethod
    def backward(ctx, dy):
        (x2,) = ctx.saved_tensors
        w = ctx.w
        dy2 = dy.reshape(-1, w.rows)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = F.gemm(dy2, w.gathered, b_major=F.MN_MAJOR).view(ctx.xshape)
        dw_full = F.gemm(dy2, x2, a_major=F.MN_MAJOR, b_major=F.MN_MAJOR)  # [rows, K] bf16
        per = w.rows // w.world
        dw_shard = torch.empty((per, w.K), dtype=torch.bfloat16, device=dy.device)
        dist.reduce_scatter_tensor(dw_shard, dw_full, op=dist.ReduceOp.AVG, group=w.group)
        db = None
        if ctx.has_bias:
            db = torch.empty_like(ctx.bias_ref)
            F.colsum(dy2, db)
            dist.all_reduce(db, op=dist.ReduceOp.AVG, group=w.group)
        return dx, dw_shard, db, None


class ShardedLinear(torch.nn.Module):
    """Drop-in for ``nn.Linear`` under ZeRO-3: this rank's parameter is only its ``out_features / world`` rows.

    The parameter's storage IS the symmetric-memory shard the peers read from inside K3, so an in-place optimizer
    update is immediately what the next forward gathers (K3's end-of-kernel flag barrier guarantees that every peer
    has finished reading a shard before its owner's stream can proceed to the update)."""

    def __init__(self, full_weight: torch.Tensor, bias: Optional[torch.Tensor] = None, group=None, device=None):
        super().__init__()
        self.sharded = ShardedWeight(full_weight, group=group, device=device)
        self.weight = torch.nn.Parameter(self.sharded.shard)  # [out/world, in], aliasing the symmetric buffer
        self.bias = None if bias is None else torch.nn.Parameter(
            bias.detach().to(self.sharded.device, torch.bfloat16).clone())
        self.in_features, self.out_features = full_weight.shape[1], full_weight.shape[0]

    def forward(self, x):
        assert self.weight.data_ptr() == self.sharded.shard.data_ptr(), "the shard must stay in symmetric memory"
        return _ShardedLinearFn.apply(x, self.weight, self.bias, self.sharded)

    def extra_repr(self):
        return f"in={self.in_features}, out={self.out_features}, rows/rank={self.out_features // self.sharded.world}"


def shard_linears(module: torch.nn.Module, group=None, device=None, min_out_features: int = 512) -> int:
    """Replace every eligible ``nn.Linear`` of ``module`` (out_features divisible by ``256 * world``) with a
    :class:`ShardedLinear`; returns how many were replaced. The remaining parameters stay replicated (wrap the module
    in :class:`maggy_b200.parallel.DataParallel` or all-reduce their gradients)."""
    world = dist.get_world_size(group)
    n = 0
    for name, child in list(module.named_children()):
        if isinstance(child, torch.nn.Linear) and child.out_features >= min_out_features and \
                child.out_features % (256 * world) == 0 and child.in_features % 8 == 0:
            setattr(module, name, ShardedLinear(child.weight.detach(), child.bias, group=group, device=device))
            n += 1
        else:
            n += shard_linears(child, group, device, min_out_features)
    return n
