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
        u64, f = ctypes.c_ulonglong, ctypes.c_float
        _lib.register_cuda_symbol("mg_ag_gemm", [P, P, P, u64p, P, P, u64p, P, i, i, i, ll, ll, i, i, i, u, P])
        _lib.register_cuda_symbol("mg_ag_gemm_mxfp8", [P, P, P, P, u64p, u64p, P, P, P, u64p, P, i, i, i, ll, ll, i, i,
                                                       i, u, P])
        _lib.register_cuda_symbol("mg_rs_sum", [u64p, u64p, u64, P, P, ll, i, i, u, f, i, P])
        _lib.register_cuda_symbol("mg_dequant_mxfp8", [P, ll, P, P, ll, i, i, P])
        _registered = True


class SymmetricReduceScatter:
    """Reduce-scatter (mean) of a full-size bf16 buffer over NVLink peer memory (csrc/comm/rs_sum.cu) — replaces
    ``dist.reduce_scatter_tensor`` in the sharded layers' backward. One workspace per (group, process): backward visits
    the layers one after the other, and the kernel's end barrier keeps a peer that is still reading from being
    overwritten by the next layer's wgrad."""

    _instances: dict = {}

    @classmethod
    def get(cls, numel: int, group=None, device=None) -> "SymmetricReduceScatter":
        key = (id(group), str(device))
        inst = cls._instances.get(key)
        if inst is None or inst.numel < numel:
            inst = cls(numel, group, device)
            cls._instances[key] = inst
        return inst

    def __init__(self, numel: int, group=None, device=None):
        import torch.distributed._symmetric_memory as symm_mem

        _register()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        gname = group.group_name if group is not None else dist.group.WORLD.group_name
        self.numel = (numel + 8 * self.world - 1) // (8 * self.world) * (8 * self.world)
        self.buf = symm_mem.empty(self.numel, dtype=torch.bfloat16, device=self.device)
        self._h = symm_mem.rendezvous(self.buf, gname)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=self.device)
        self.flags.zero_()
        self._hf = symm_mem.rendezvous(self.flags, gname)
        self.local_sync = torch.zeros(4, dtype=torch.int32, device=self.device)
        arr = ctypes.c_ulonglong * self.world
        self._in_ptrs = arr(*[int(p) for p in self._h.buffer_ptrs])
        self._flag_ptrs = arr(*[int(p) for p in self._hf.buffer_ptrs])
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def view(self, rows: int, cols: int) -> torch.Tensor:
        """The workspace as a [rows, cols] matrix: have the wgrad GEMM write straight into it."""
        assert rows * cols <= self.numel
        return self.buf[:rows * cols].view(rows, cols)

    def reduce_scatter_mean(self, rows: int, cols: int, out: torch.Tensor) -> torch.Tensor:
        """``out`` [rows / world, cols] = mean over ranks of this rank's row block of the [rows, cols] workspace."""
        shard = rows * cols // self.world
        assert shard % 8 == 0 and out.numel() == shard and out.is_contiguous() and out.dtype == torch.bfloat16
        self.epoch += 1
        # the kernel addresses shard r at element r * shard_elems of every peer's buffer
        rc = _lib.cuda().mg_rs_sum(self._in_ptrs, self._flag_ptrs, 0, out.data_ptr(), self.local_sync.data_ptr(),
                                   shard, self.rank, self.world, self.epoch, 1.0 / self.world, 0, F._stream())
        _lib.check(rc, "mg_rs_sum")
        F._count(1)
        return out


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
    backward : dgrad on the gathered copy (tcgen05 GEMM, MN-major B), full wgrad GEMM written into a symmetric
               workspace, then the peer-memory reduce-scatter kernel (csrc/comm/rs_sum.cu) so that each rank keeps only
               the gradient of its own rows (averaged over the group).
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
        # full wgrad straight into the symmetric workspace, then each rank sums ITS rows out of every peer's copy
        # (csrc/comm/rs_sum.cu: peer loads over NVLink; no NCCL reduce-scatter)
        per = w.rows // w.world
        rs = SymmetricReduceScatter.get(w.rows * w.K, w.group, dy.device)
        F.gemm(dy2, x2, a_major=F.MN_MAJOR, b_major=F.MN_MAJOR, out=rs.view(w.rows, w.K))  # [rows, K] bf16
        dw_shard = rs.reduce_scatter_mean(w.rows, w.K, torch.empty((per, w.K), dtype=torch.bfloat16, device=dy.device))
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


# ------------------------------------------------------------------------------------------------
# block-scaled fp8 (MXFP8): the shards travel as e4m3 bytes + UE8M0 scales, half the NVLink bytes of bf16
# ------------------------------------------------------------------------------------------------
class ShardedWeightMX:
    """Rows ``[rank*rows/world, ...)`` of a ``[rows, K]`` weight: a bf16 master shard (the trainable parameter) plus
    its MXFP8 image (e4m3 bytes + scale chunks) in symmetric memory, re-quantised from the master before every forward.
    Peers gather the fp8 image inside the GEMM kernel (csrc/comm/ag_gemm_mxfp8.cu)."""

    def __init__(self, full_weight: torch.Tensor, group=None, device=None, n_copy: int = 0):
        import torch.distributed._symmetric_memory as symm_mem

        _register()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        rows, K = full_weight.shape
        assert rows % self.world == 0 and (rows // self.world) % 128 == 0, "rows per shard must be a multiple of 128"
        assert K % 16 == 0
        self.rows, self.K = rows, K
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        gname = group.group_name if group is not None else dist.group.WORLD.group_name
        per = rows // self.world
        self.per = per
        self.master = full_weight[self.rank * per:(self.rank + 1) * per].to(self.device, torch.bfloat16).contiguous()
        nkb = (K + 127) // 128
        self.sf_shard_bytes = (per // 128) * nkb * 512
        self.q_shard = symm_mem.empty((per, K), dtype=torch.uint8, device=self.device)
        self._h_q = symm_mem.rendezvous(self.q_shard, gname)
        self.sf_shard = symm_mem.empty(self.sf_shard_bytes, dtype=torch.uint8, device=self.device)
        self._h_sf = symm_mem.rendezvous(self.sf_shard, gname)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=self.device)
        self.flags.zero_()
        self._h_flags = symm_mem.rendezvous(self.flags, gname)
        self.gathered_q = torch.empty((rows, K), dtype=torch.uint8, device=self.device)
        self.gathered_sf = torch.empty((rows // 128) * nkb * 512, dtype=torch.uint8, device=self.device)
        self.ready = torch.zeros(rows // 128, dtype=torch.int32, device=self.device)
        self.local_sync = torch.zeros(4, dtype=torch.int32, device=self.device)
        arr = ctypes.c_ulonglong * self.world
        self._q_ptrs = arr(*[int(p) for p in self._h_q.buffer_ptrs])
        self._sf_ptrs = arr(*[int(p) for p in self._h_sf.buffer_ptrs])
        self._flag_ptrs = arr(*[int(p) for p in self._h_flags.buffer_ptrs])
        self.n_copy = n_copy
        self.epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def requantize(self, master: torch.Tensor):
        """bf16 master shard -> the fp8 image the peers read (same stream as the GEMM that follows)."""
        lib = _lib.cuda()
        rc = lib.mg_quant_mxfp8(master.data_ptr(), master.stride(0), self.q_shard.data_ptr(), self.K,
                                self.sf_shard.data_ptr(), self.per, self.K, F._stream())
        _lib.check(rc, "mg_quant_mxfp8(shard)")
        F._count(1)

    def gathered_bf16(self) -> torch.Tensor:
        """The full weight in bf16, dequantised from what the last forward gathered (+ the local shard's image)."""
        per, K = self.per, self.K
        nkb = (K + 127) // 128
        lo = self.rank * per
        self.gathered_q[lo:lo + per].copy_(self.q_shard)
        self.gathered_sf[(lo // 128) * nkb * 512:((lo + per) // 128) * nkb * 512].copy_(self.sf_shard)
        out = torch.empty((self.rows, K), dtype=torch.bfloat16, device=self.device)
        rc = _lib.cuda().mg_dequant_mxfp8(self.gathered_q.data_ptr(), K, self.gathered_sf.data_ptr(), out.data_ptr(), K,
                                          self.rows, K, F._stream())
        _lib.check(rc, "mg_dequant_mxfp8")
        F._count(1)
        return out


def ag_linear_mx(x: torch.Tensor, w: ShardedWeightMX, bias: Optional[torch.Tensor] = None, out=None) -> torch.Tensor:
    """``x[..., K] @ W_fullᵀ (+ bias)`` in block-scaled fp8 with the all-gather of W's fp8 image fused into the GEMM."""
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, K = x2.shape
    assert K == w.K and x2.dtype == torch.bfloat16
    xq, sfx = F.quantize_mxfp8(x2)
    if out is None:
        out = torch.empty((M, w.rows), dtype=torch.bfloat16, device=x.device)
    w.epoch += 1
    rc = _lib.cuda().mg_ag_gemm_mxfp8(xq.data_ptr(), sfx.data_ptr(), out.data_ptr(), 0 if bias is None else bias.data_ptr(),
                                      w._q_ptrs, w._sf_ptrs, w.gathered_q.data_ptr(), w.gathered_sf.data_ptr(),
                                      w.ready.data_ptr(), w._flag_ptrs, w.local_sync.data_ptr(), M, w.rows, K,
                                      xq.stride(0), out.stride(0), w.rank, w.world, w.n_copy, w.epoch, F._stream())
    _lib.check(rc, "mg_ag_gemm_mxfp8")
    F._count(1)
    return out.view(*x.shape[:-1], w.rows)


class _ShardedLinearMXFn(torch.autograd.Function):
    """forward: requantise the local shard, K3-MX (gather of fp8 images fused into the block-scaled GEMM);
    backward: bf16 dgrad on the dequantised gathered weight, bf16 wgrad + peer-memory reduce-scatter."""

    @staticmethod
    def forward(ctx, x, master, bias, w: ShardedWeightMX):
        w.requantize(master.detach())
        ctx.w, ctx.has_bias, ctx.bias_ref = w, bias is not None, bias
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ctx.save_for_backward(x2)
        ctx.xshape = x.shape
        return ag_linear_mx(x2, w, bias).view(*x.shape[:-1], w.rows)

    @staticmethod
    def backward(ctx, dy):
        (x2,) = ctx.saved_tensors
        w = ctx.w
        dy2 = dy.reshape(-1, w.rows)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = F.gemm(dy2, w.gathered_bf16(), b_major=F.MN_MAJOR).view(ctx.xshape)
        rs = SymmetricReduceScatter.get(w.rows * w.K, w.group, dy.device)
        F.gemm(dy2, x2, a_major=F.MN_MAJOR, b_major=F.MN_MAJOR, out=rs.view(w.rows, w.K))
        dw_shard = rs.reduce_scatter_mean(w.rows, w.K, torch.empty((w.per, w.K), dtype=torch.bfloat16, device=dy.device))
        db = None
        if ctx.has_bias:
            db = torch.empty_like(ctx.bias_ref)
            F.colsum(dy2, db)
            dist.all_reduce(db, op=dist.ReduceOp.AVG, group=w.group)
        return dx, dw_shard, db, None


class ShardedLinearMX(torch.nn.Module):
    """``nn.Linear`` under ZeRO-3 whose forward runs in block-scaled fp8: the parameter is this rank's bf16 row shard;
    what crosses NVLink is its MXFP8 image."""

    def __init__(self, full_weight: torch.Tensor, bias: Optional[torch.Tensor] = None, group=None, device=None):
        super().__init__()
        self.sharded = ShardedWeightMX(full_weight, group=group, device=device)
        self.weight = torch.nn.Parameter(self.sharded.master)
        self.bias = None if bias is None else torch.nn.Parameter(
            bias.detach().to(self.sharded.device, torch.bfloat16).clone())
        self.in_features, self.out_features = full_weight.shape[1], full_weight.shape[0]

    def forward(self, x):
        return _ShardedLinearMXFn.apply(x, self.weight, self.bias, self.sharded)

    def extra_repr(self):
        return (f"in={self.in_features}, out={self.out_features}, rows/rank={self.sharded.per}, "
                "compute=mxfp8 (e4m3 + UE8M0 / 32)")


def shard_linears(module: torch.nn.Module, group=None, device=None, min_out_features: int = 512,
                  fp8: bool = False) -> int:
    """Replace every eligible ``nn.Linear`` of ``module`` (out_features divisible by ``256 * world``) with a
    :class:`ShardedLinear` (or, with ``fp8=True``, a block-scaled :class:`ShardedLinearMX`); returns how many were
    replaced. The remaining parameters stay replicated (wrap the module in :class:`maggy_b200.parallel.DataParallel` or
    all-reduce their gradients)."""
    world = dist.get_world_size(group)
    n = 0
    for name, child in list(module.named_children()):
        if isinstance(child, torch.nn.Linear) and child.out_features >= min_out_features and \
                child.out_features % (256 * world) == 0 and child.in_features % 16 == 0:
            cls = ShardedLinearMX if fp8 else ShardedLinear
            setattr(module, name, cls(child.weight.detach(), child.bias, group=group, device=device))
            n += 1
        else:
            n += shard_linears(child, group, device, min_out_features, fp8)
    return n
