"""Python front-end of the hand-written sm_100a kernels (``csrc/ops``).

Two layers:

* thin launchers (``gemm``, ``layer_norm_fwd`` ...) that validate shapes/strides and call the C ABI of
  ``libmaggy_b200_cuda.so`` on the current torch CUDA stream (so they are CUDA-graph capturable);
* ``torch.autograd.Function`` wrappers (``linear``, ``layer_norm``, ``mlp``, ``cross_entropy``,
  ``embedding``) whose backward also runs on the hand-written kernels.

Weight gradients follow the *main_grad* convention: if a parameter has a ``main_grad`` attribute (a
view into the flat gradient arena owned by ``maggy_b200.parallel``), wgrad kernels write straight
into it and autograd sees ``None`` — no per-parameter ``.grad`` tensors, no accumulate kernels, and
the arena is exactly the buffer the fused reduce-scatter+Adam kernel consumes.

These ops require CUDA: calling them on CPU tensors raises (tests use plain PyTorch as the fp32
reference, never as a fallback).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_BIAS_RESIDUAL = 0, 1, 2, 3, 4
# Column sums of d_pre inside the dGELU dgrad epilogue: correct (tests) but measured SLOWER end to end on the same box
# (GPT-2 small, B = 64: 990.3 k / 984.2 k tokens/s without vs 979.0 k / 975.9 k with) — that GEMM's epilogue is already
# its critical path and the read-back adds 128 LDS + 256 FADD per thread per tile. Off by default, kept for the record.
_DGELU_COLSUM = __import__("os").environ.get("MAGGY_B200_DGELU_COLSUM", "0") != "0"
K_MAJOR, MN_MAJOR = 0, 1

_scratch: dict = {}
_launches = 0  # hand-written kernels launched through this module (bench.py reports it)


def launch_count() -> int:
    return _launches


def _count(n: int = 1):
    global _launches
    _launches += n


def _stream() -> int:
    """Current torch stream handle; also binds the kernel library's runtime to torch's current device on this
    thread (autograd worker threads start without a bound driver context)."""
    _lib.cuda().mg_set_device(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("maggy_b200 ops run on CUDA tensors only (sm_100a kernels); got a CPU tensor")


def _scratch_f32(key: str, numel: int, device) -> torch.Tensor:
    """Reusable fp32 workspace (stable address => CUDA-graph safe as long as sizes are stable)."""
    k = (key, device)
    buf = _scratch.get(k)
    if buf is None or buf.numel() < numel:
        buf = torch.empty(max(numel, 1), dtype=torch.float32, device=device)
        _scratch[k] = buf
    return buf


# ------------------------------------------------------------------------------------------------
# thin launchers
# ------------------------------------------------------------------------------------------------
def gemm(a, b, *, a_major=K_MAJOR, b_major=K_MAJOR, bias=None, aux=None, epi=EPI_NONE, out=None,
         out_dtype=torch.bfloat16, accumulate=False, block_n=0, sm_limit=0, splits=None):
    """``C[M,N] = epi(sum_k A(m,k) B(n,k))`` on the tcgen05 kernel.

    ``a`` is ``[M,K]`` (``a_major=K_MAJOR``) or ``[K,M]`` (``MN_MAJOR``); ``b`` is ``[N,K]`` or ``[K,N]``.
    Rows may be strided (``stride(0)`` multiple of 8 elements), the inner dimension must be contiguous.
    ``splits``: split-K factor; ``None`` picks one automatically for small outputs with a long reduction
    (weight gradients), ``1`` disables it.
    """
    _need_cuda(a, b, bias, aux, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, "bf16 operands"
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[0], a.shape[1]) if a_major == K_MAJOR else (a.shape[1], a.shape[0])
    N, Kb = (b.shape[0], b.shape[1]) if b_major == K_MAJOR else (b.shape[1], b.shape[0])
    assert K == Kb, f"reduction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    assert out.dtype in (torch.bfloat16, torch.float32)
    ld_aux = 0
    if epi in (EPI_BIAS_GELU, EPI_DGELU, EPI_BIAS_RESIDUAL):
        assert aux is not None and aux.shape == (M, N) and aux.stride(1) == 1 and aux.dtype == torch.bfloat16
        ld_aux = aux.stride(0)
    if bias is not None and epi == EPI_DGELU:
        # not a bias: the fp32 buffer that receives the column sums of the output per 32-row slice (see _MLP.backward)
        assert bias.dtype == torch.float32 and bias.numel() >= 4 * ((M + 127) // 128) * N and bias.is_contiguous()
        assert out.dtype == torch.bfloat16 and not accumulate
    elif bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.numel() == N and bias.is_contiguous()
    lib = _lib.cuda()
    stream = _stream()
    ws = None
    if splits is None:
        splits = 1
        if epi == EPI_NONE and out.stride(0) == N:
            splits = lib.mg_gemm_auto_splits(M, N, K, block_n)
    if splits > 1:
        assert epi == EPI_NONE and out.stride(0) == N, "split-K needs a plain epilogue and a dense output"
        ws = _scratch_f32("splitk", splits * M * N, a.device)
    rc = lib.mg_gemm_bf16(
        a.data_ptr(), b.data_ptr(), out.data_ptr(), _ptr(bias), _ptr(aux), M, N, K,
        a.stride(0), b.stride(0), out.stride(0), ld_aux, a_major, b_major, epi,
        1 if out.dtype == torch.float32 else 0, 1 if accumulate else 0, block_n, sm_limit, splits,
        _ptr(ws), stream)
    _lib.check(rc, f"mg_gemm_bf16(M={M},N={N},K={K})")
    _count(2 if splits > 1 else 1)
    return out


def layer_norm_fwd(x2d, weight, bias, eps=1e-5, rms=False):
    _need_cuda(x2d, weight, bias)
    rows, C = x2d.shape
    assert x2d.is_contiguous() and x2d.dtype == torch.bfloat16
    y = torch.empty_like(x2d)
    mean = None if rms else torch.empty(rows, dtype=torch.float32, device=x2d.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    rc = _lib.cuda().mg_norm_fwd(x2d.data_ptr(), weight.data_ptr(), _ptr(bias), y.data_ptr(),
                                 _ptr(mean), rstd.data_ptr(), rows, C, eps, 1 if rms else 0, _stream())
    _lib.check(rc, "mg_norm_fwd")
    _count(1)
    return y, mean, rstd


def layer_norm_bwd(dy2d, x2d, weight, mean, rstd, dgamma, dbeta, rms=False, accumulate=False, res_grad=None,
                   dcol=None, accumulate_col=False):
    """``res_grad`` ([rows, C] bf16, optional) is added to the returned ``dx`` inside the kernel (the gradient that
    reaches ``x`` through a residual connection around the normalised branch). ``dcol`` ([C] bf16, optional) receives
    the column sums of that ``dx`` — the bias gradient of the layer that produced ``x`` — from the same pass."""
    _need_cuda(dy2d, x2d, res_grad)
    rows, C = x2d.shape
    dy2d = dy2d.contiguous()
    if res_grad is not None:
        res_grad = res_grad.reshape(rows, C)
        if not res_grad.is_contiguous():
            res_grad = res_grad.contiguous()
    dx = torch.empty_like(x2d)
    lib = _lib.cuda()
    grid = lib.mg_norm_bwd_grid(rows)
    partial = _scratch_f32("norm_bwd", grid * 2 * C, x2d.device)
    fused_col = dcol is not None and bool(lib.mg_norm_bwd_supports_col(C))
    rc = lib.mg_norm_bwd_col(dy2d.data_ptr(), x2d.data_ptr(), weight.data_ptr(), _ptr(mean),
                             rstd.data_ptr(), dx.data_ptr(), dgamma.data_ptr(), _ptr(dbeta),
                             dcol.data_ptr() if fused_col else None, partial.data_ptr(), rows, C, 1 if rms else 0,
                             1 if accumulate else 0, 1 if accumulate_col else 0, _ptr(res_grad), _stream())
    _lib.check(rc, "mg_norm_bwd")
    _count(2)
    if dcol is not None and not fused_col:  # widths the staged kernel does not cover: a separate pass over dx
        colsum(dx, dcol, accumulate=accumulate_col)
    return dx


def colsum(x2d, out, accumulate=False):
    """out[n] (+)= sum_m x2d[m, n] — bias gradients."""
    _need_cuda(x2d, out)
    M, N = x2d.shape
    assert x2d.stride(1) == 1 and out.numel() == N
    lib = _lib.cuda()
    partial = _scratch_f32("colsum", lib.mg_colsum_slabs(M) * N, x2d.device)
    rc = lib.mg_colsum_bf16(x2d.data_ptr(), M, N, x2d.stride(0), out.data_ptr(), partial.data_ptr(),
                            1 if accumulate else 0, _stream())
    _lib.check(rc, "mg_colsum_bf16")
    _count(2)
    return out


def gelu_fwd(x):
    _need_cuda(x)
    y = torch.empty_like(x)
    _lib.check(_lib.cuda().mg_gelu_fwd(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "mg_gelu_fwd")
    _count(1)
    return y


def gelu_bwd(x, dy):
    _need_cuda(x, dy)
    dx = torch.empty_like(x)
    _lib.check(_lib.cuda().mg_gelu_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), _stream()),
               "mg_gelu_bwd")
    _count(1)
    return dx


def adam_flat_step(master, m, v, grad_bf16, param_bf16, *, decay_boundary, lr, beta1, beta2, eps,
                   weight_decay, step, grad_scale=1.0, gnorm_sq=None, max_norm=0.0, step_tensor=None,
                   lr_tensor=None):
    _need_cuda(master, m, v, grad_bf16, param_bf16)
    n = master.numel()
    rc = _lib.cuda().mg_adam_flat(master.data_ptr(), m.data_ptr(), v.data_ptr(), grad_bf16.data_ptr(),
                                  param_bf16.data_ptr(), n, decay_boundary, lr, beta1, beta2, eps,
                                  weight_decay, step, grad_scale, _ptr(gnorm_sq), max_norm,
                                  _ptr(step_tensor), _ptr(lr_tensor), _stream())
    _lib.check(rc, "mg_adam_flat")
    _count(1)


def sumsq(x_bf16, out_f32):
    _need_cuda(x_bf16, out_f32)
    out_f32.zero_()
    _lib.check(_lib.cuda().mg_sumsq_bf16(x_bf16.data_ptr(), x_bf16.numel(), out_f32.data_ptr(), _stream()),
               "mg_sumsq_bf16")
    _count(1)
    return out_f32


def quantize_fp8(x: torch.Tensor):
    """Per-tensor e4m3 quantisation on the device: returns ``(q, scale)`` with ``q`` a ``torch.float8_e4m3fn``
    tensor of ``x``'s shape and ``scale`` a 1-element fp32 device tensor such that ``x ~= q * scale``."""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    ws = torch.empty(2, dtype=torch.float32, device=x.device)
    lib = _lib.cuda()
    rc = lib.mg_quant_e4m3(x.data_ptr(), q.data_ptr(), x.numel(), ws.data_ptr(), ws[1:].data_ptr(), _stream())
    _lib.check(rc, "mg_quant_e4m3")
    _count(2)
    return q.view(torch.float8_e4m3fn), ws[1:]


def gemm_fp8(a_q, b_q, scale_a=None, scale_b=None, *, bias=None, out=None, alpha: float = 1.0, sm_limit=0):
    """``C[M,N] (bf16) = alpha * scale_a * scale_b * sum_k A(m,k) B(n,k) (+ bias)`` with e4m3 operands
    (``a_q`` [M,K], ``b_q`` [N,K], both K-major) on ``tcgen05.mma.kind::f8f6f4``; the scales are device scalars."""
    _need_cuda(a_q, b_q, bias, out)
    assert a_q.element_size() == 1 and b_q.element_size() == 1 and a_q.dim() == 2 and b_q.dim() == 2
    assert a_q.stride(1) == 1 and b_q.stride(1) == 1
    M, K = a_q.shape
    N, Kb = b_q.shape
    assert K == Kb, f"reduction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a_q.device)
    assert out.dtype == torch.bfloat16 and out.shape == (M, N) and out.stride(1) == 1
    lib = _lib.cuda()
    rc = lib.mg_gemm_fp8(a_q.data_ptr(), b_q.data_ptr(), out.data_ptr(), _ptr(bias), M, N, K, a_q.stride(0),
                         b_q.stride(0), out.stride(0), float(alpha), _ptr(scale_a), _ptr(scale_b),
                         EPI_BIAS if bias is not None else EPI_NONE, sm_limit, _stream())
    _lib.check(rc, f"mg_gemm_fp8(M={M},N={N},K={K})")
    _count(1)
    return out


def attention_fwd(qkv: torch.Tensor, n_heads: int, n_kv_heads: Optional[int] = None, causal: bool = True,
                  scale: Optional[float] = None):
    """Fused attention forward on tcgen05 (head_dim 64). ``qkv``: [B, T, (H + 2*KVH) * 64] bf16, contiguous — the
    fused projection output, q heads first, then k heads, then v heads. Returns ``(o [B, T, H*64], lse [B, H, T])``
    (``lse`` in the log2 domain of the scaled scores)."""
    _need_cuda(qkv)
    B, T, Cw = qkv.shape
    kvh = n_heads if n_kv_heads is None else n_kv_heads
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and Cw == (n_heads + 2 * kvh) * 64
    o = torch.empty((B, T, n_heads * 64), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, n_heads, T), dtype=torch.float32, device=qkv.device)
    sc = 0.125 if scale is None else float(scale)
    rc = _lib.cuda().mg_attn_fwd(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, n_heads, kvh,
                                 1 if causal else 0, sc, _stream())
    _lib.check(rc, f"mg_attn_fwd(B={B},T={T},H={n_heads},KVH={kvh})")
    _count(1)
    return o, lse


class _Attention(torch.autograd.Function):
    """Fused attention over the packed projection output (see :func:`attention`)."""

    @staticmethod
    def forward(ctx, qkv, n_heads, n_kv_heads, causal, scale):
        if not qkv.is_contiguous():
            qkv = qkv.contiguous()
        o, lse = attention_fwd(qkv, n_heads, n_kv_heads, causal, scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.cfg = (n_heads, n_kv_heads, causal, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        H, KVH, causal, scale = ctx.cfg
        B, T, W = qkv.shape
        if not do.is_contiguous():
            do = do.contiguous()
        dqkv = torch.empty_like(qkv)
        delta = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
        dq_acc = torch.empty((B, T, H * 64), dtype=torch.float32, device=qkv.device)
        sc = 0.125 if scale is None else float(scale)
        rc = _lib.cuda().mg_attn_bwd(qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                     delta.data_ptr(), dq_acc.data_ptr(), B, T, H, KVH, 1 if causal else 0, sc,
                                     _stream())
        _lib.check(rc, f"mg_attn_bwd(B={B},T={T},H={H},KVH={KVH})")
        _count(4)
        return dqkv, None, None, None, None


def attention(qkv: torch.Tensor, n_heads: int, n_kv_heads: Optional[int] = None, causal: bool = True,
              scale: Optional[float] = None) -> torch.Tensor:
    """``softmax(scale * Q K^T [+ causal mask]) V`` on the tcgen05 attention kernels, head_dim 64.

    ``qkv`` is the fused projection output ``[B, T, (H + 2*KVH) * 64]`` (q heads, then k heads, then v heads);
    the result is ``[B, T, H*64]``, i.e. already in the layout of the output projection's input. Backward returns
    the gradient in the same packed layout, so no split / transpose / concatenation kernels run in either direction."""
    kvh = n_heads if n_kv_heads is None else n_kv_heads
    return _Attention.apply(qkv, n_heads, kvh, causal, scale)


def _write_or_return_grad(param, compute_into):
    """main_grad convention. ``compute_into(dst, accumulate)`` must fill ``dst``."""
    mg = getattr(param, "main_grad", None)
    if mg is not None:
        acc = bool(getattr(param, "_mg_written", False))
        compute_into(mg, acc)
        param._mg_written = True
        return None
    dst = torch.empty_like(param)
    compute_into(dst, False)
    return dst


# ------------------------------------------------------------------------------------------------
# autograd functions
# ------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    """y = x W^T + b with tcgen05 fprop / dgrad / wgrad. Optional fused residual add."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        if residual is not None:
            r2 = residual.reshape(-1, weight.shape[0])
            y = gemm(x2, weight, bias=bias, aux=r2, epi=EPI_BIAS_RESIDUAL)
        else:
            y = gemm(x2, weight, bias=bias, epi=EPI_BIAS if bias is not None else EPI_NONE)
        ctx.save_for_backward(x2, weight)
        ctx.bias = bias
        ctx.weight_ref = weight
        ctx.has_res = residual is not None
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy2, weight, b_major=MN_MAJOR).view(ctx.xshape)
        dw = _write_or_return_grad(
            ctx.weight_ref,
            lambda dst, acc: gemm(dy2, x2, a_major=MN_MAJOR, b_major=MN_MAJOR, out=dst, accumulate=acc))
        db = None
        if ctx.bias is not None and not _take_norm_colsum(ctx.bias):
            db = _write_or_return_grad(ctx.bias, lambda dst, acc: colsum(dy2, dst, accumulate=acc))
        dres = dy if ctx.has_res else None
        return dx, dw, db, dres


def linear(x, weight, bias=None, residual=None):
    return _Linear.apply(x, weight, bias, residual)


class _LinearFP8(torch.autograd.Function):
    """y = x W^T (+ b) with the forward GEMM in e4m3 (per-tensor dynamic scales, quantised on the device);
    dgrad / wgrad stay in bf16 on the saved bf16 operands."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        xq, sx = quantize_fp8(x2)
        wq, sw = quantize_fp8(weight if weight.is_contiguous() else weight.contiguous())
        y = gemm_fp8(xq, wq, sx, sw, bias=bias)
        ctx.save_for_backward(x2, weight)
        ctx.bias, ctx.weight_ref, ctx.xshape = bias, weight, x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = gemm(dy2, weight, b_major=MN_MAJOR).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = _write_or_return_grad(
            ctx.weight_ref,
            lambda dst, acc: gemm(dy2, x2, a_major=MN_MAJOR, b_major=MN_MAJOR, out=dst, accumulate=acc))
        db = None
        if ctx.bias is not None:
            db = _write_or_return_grad(ctx.bias, lambda dst, acc: colsum(dy2, dst, accumulate=acc))
        return dx, dw, db


def linear_fp8(x, weight, bias=None):
    return _LinearFP8.apply(x, weight, bias)


def quantize_mxfp8(x: torch.Tensor):
    """Block-scaled (OCP MXFP8) quantisation on the device: ``(q, sf)`` with ``q`` a ``torch.float8_e4m3fn`` tensor of
    ``x``'s shape [rows, K] and ``sf`` the UE8M0 scale bytes — one per 32 elements along K — already in the chunked
    layout the GEMM's tcgen05.cp consumes (csrc/ops/gemm_mxfp8.cu). ``K`` must be a multiple of 16."""
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 16 == 0
    rows, K = x.shape
    lib = _lib.cuda()
    q = torch.empty((rows, K), dtype=torch.uint8, device=x.device)
    sf = torch.empty(int(lib.mg_mxfp8_sf_bytes(rows, K)), dtype=torch.uint8, device=x.device)
    rc = lib.mg_quant_mxfp8(x.data_ptr(), x.stride(0), q.data_ptr(), q.stride(0), sf.data_ptr(), rows, K, _stream())
    _lib.check(rc, f"mg_quant_mxfp8(rows={rows},K={K})")
    _count(1)
    return q.view(torch.float8_e4m3fn), sf


def dequantize_mxfp8(q: torch.Tensor, sf: torch.Tensor) -> torch.Tensor:
    """fp32 values of a block-scaled operand (reference / test helper; plain PyTorch)."""
    rows, K = q.shape
    nkb = (K + 127) // 128
    r = torch.arange(rows, device=q.device).view(-1, 1)
    kb32 = (torch.arange(K, device=q.device) // 32).view(1, -1)
    idx = ((r // 128) * nkb + kb32 // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + kb32 % 4
    e = sf.to(torch.int32)[idx].float() - 127.0
    return q.float() * torch.exp2(e)


def gemm_mxfp8(a_q, sfa, b_q, sfb, *, bias=None, out=None, sm_limit=0):
    """``C[M,N] (bf16) = A_mx . B_mx^T (+ bias)`` on ``tcgen05.mma.kind::mxf8f6f4.block_scale``: e4m3 operands
    (``a_q`` [M,K], ``b_q`` [N,K], K-major) with one UE8M0 scale per 32 K-elements applied by the tensor core."""
    _need_cuda(a_q, b_q, bias, out)
    assert a_q.element_size() == 1 and b_q.element_size() == 1 and a_q.stride(1) == 1 and b_q.stride(1) == 1
    M, K = a_q.shape
    N, Kb = b_q.shape
    assert K == Kb, f"reduction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a_q.device)
    rc = _lib.cuda().mg_gemm_mxfp8(a_q.data_ptr(), sfa.data_ptr(), b_q.data_ptr(), sfb.data_ptr(), out.data_ptr(),
                                   _ptr(bias), M, N, K, a_q.stride(0), b_q.stride(0), out.stride(0), sm_limit, _stream())
    _lib.check(rc, f"mg_gemm_mxfp8(M={M},N={N},K={K})")
    _count(1)
    return out


class _LinearMXFP8(torch.autograd.Function):
    """y = x W^T (+ b) with the forward GEMM in block-scaled fp8 (MXFP8: e4m3 + UE8M0 per 32 elements, quantised on
    the device); dgrad / wgrad stay in bf16 on the saved bf16 operands."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        xq, sfx = quantize_mxfp8(x2)
        wq, sfw = quantize_mxfp8(weight if weight.is_contiguous() else weight.contiguous())
        y = gemm_mxfp8(xq, sfx, wq, sfw, bias=bias)
        ctx.save_for_backward(x2, weight)
        ctx.bias, ctx.weight_ref, ctx.xshape = bias, weight, x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    backward = _LinearFP8.backward


def linear_mxfp8(x, weight, bias=None):
    return _LinearMXFP8.apply(x, weight, bias)


def _take_norm_colsum(bias) -> bool:
    """True when the norm that consumed this layer's output already produced ``bias``'s gradient (``col_bias=``)."""
    if bias is not None and getattr(bias, "_colsum_in_norm", False):
        bias._colsum_in_norm = False
        return True
    return False


class _LayerNorm(torch.autograd.Function):
    """``with_residual=True`` additionally returns ``x`` itself (an alias, no kernel) as the tensor to feed into the
    residual connection around the normalised branch: both gradients then arrive in THIS backward and are summed
    inside the norm-backward kernel instead of by a separate autograd add over the whole activation.

    ``col_bias``: the bias parameter of the ``linear`` / ``mlp`` whose output is ``x`` (and is consumed ONLY here).
    The ``dx`` this backward writes is that layer's output gradient, so its bias gradient (column sums of ``dx``) is
    accumulated by the same kernel and the producer skips its own reduction pass."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, rms, with_residual=False, col_bias=None):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y, mean, rstd = layer_norm_fwd(x2, weight, bias, eps, rms)
        ctx.save_for_backward(x2, weight, mean if mean is not None else rstd, rstd)
        ctx.w_ref, ctx.b_ref, ctx.rms = weight, bias, rms
        ctx.xshape = x.shape
        ctx.with_residual = with_residual
        ctx.col_bias = None
        if col_bias is not None and ctx.needs_input_grad[0]:  # (grad mode is off inside forward: ask ctx)
            ctx.col_bias = col_bias
            col_bias._colsum_in_norm = True  # consumed (and cleared) by the producer's backward
        if with_residual:
            return y.view(x.shape), x.view_as(x)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy, dres=None):
        x2, weight, mean, rstd = ctx.saved_tensors
        if dy is None:  # only the residual stream carried a gradient
            dy = torch.zeros(ctx.xshape, dtype=x2.dtype, device=x2.device)
        dy2 = dy.reshape(-1, dy.shape[-1])
        w, b = ctx.w_ref, ctx.b_ref
        use_main = getattr(w, "main_grad", None) is not None
        if use_main:
            acc = bool(getattr(w, "_mg_written", False))
            dgamma = w.main_grad
            dbeta = b.main_grad if b is not None else None
            w._mg_written = True
            if b is not None:
                b._mg_written = True
        else:
            acc = False
            dgamma = torch.empty_like(w)
            dbeta = torch.empty_like(b) if b is not None else None
        cb, dcol, acc_col, dcol_ret = ctx.col_bias, None, False, None
        if cb is not None:
            if getattr(cb, "main_grad", None) is not None:
                dcol, acc_col = cb.main_grad, bool(getattr(cb, "_mg_written", False))
                cb._mg_written = True
            else:
                dcol = dcol_ret = torch.empty_like(cb)
        dx = layer_norm_bwd(dy2, x2, weight, None if ctx.rms else mean, rstd, dgamma, dbeta,
                            rms=ctx.rms, accumulate=acc, res_grad=dres if ctx.with_residual else None,
                            dcol=dcol, accumulate_col=acc_col)
        if use_main:
            return dx.view(ctx.xshape), None, None, None, None, None, dcol_ret
        return dx.view(ctx.xshape), dgamma, dbeta, None, None, None, dcol_ret


def layer_norm(x, weight, bias, eps=1e-5, with_residual=False, col_bias=None):
    """``with_residual=True`` returns ``(norm(x), x_res)``: use ``x_res`` as the residual input of the branch
    (``linear(..., residual=x_res)`` / ``mlp(..., residual=x_res)``) to fold the residual-gradient add into the norm
    backward kernel. ``col_bias``: see :class:`_LayerNorm`."""
    return _LayerNorm.apply(x, weight, bias, eps, False, with_residual, col_bias)


def rms_norm(x, weight, eps=1e-6, with_residual=False, col_bias=None):
    return _LayerNorm.apply(x, weight, None, eps, True, with_residual, col_bias)


class _MLP(torch.autograd.Function):
    """out = residual + W2 gelu(W1 x + b1) + b2, GELU/dGELU/residual fused into GEMM epilogues."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        pre = torch.empty((M, w1.shape[0]), dtype=torch.bfloat16, device=x.device)
        act = gemm(x2, w1, bias=b1, aux=pre, epi=EPI_BIAS_GELU)
        if residual is not None:
            r2 = residual.reshape(-1, w2.shape[0])
            out = gemm(act, w2, bias=b2, aux=r2, epi=EPI_BIAS_RESIDUAL)
        else:
            out = gemm(act, w2, bias=b2, epi=EPI_BIAS)
        ctx.save_for_backward(x2, pre, act)
        ctx.refs = (w1, b1, w2, b2)
        ctx.has_res = residual is not None
        ctx.xshape = x.shape
        return out.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, pre, act = ctx.saved_tensors
        w1, b1, w2, b2 = ctx.refs
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        # d_pre = (dy W2) * gelu'(pre): dGELU fused into the dgrad epilogue — which also leaves the column sums of d_pre
        # (= the fc bias gradient) behind as fp32 partial rows, one per 32-row warp slice of the output
        M_, N1 = dy2.shape[0], w2.shape[1]
        fuse_db1 = b1 is not None and _DGELU_COLSUM and N1 % 8 == 0
        cs_rows = 4 * ((M_ + 127) // 128)
        cs_part = _scratch_f32("dgelu_colsum", cs_rows * N1, dy2.device) if fuse_db1 else None
        dpre = gemm(dy2, w2, b_major=MN_MAJOR, aux=pre, epi=EPI_DGELU, bias=cs_part)
        dw2 = _write_or_return_grad(
            w2, lambda dst, acc: gemm(dy2, act, a_major=MN_MAJOR, b_major=MN_MAJOR, out=dst, accumulate=acc))
        db2 = None
        if not _take_norm_colsum(b2):
            db2 = _write_or_return_grad(b2, lambda dst, acc: colsum(dy2, dst, accumulate=acc))
        dx = gemm(dpre, w1, b_major=MN_MAJOR).view(ctx.xshape)
        dw1 = _write_or_return_grad(
            w1, lambda dst, acc: gemm(dpre, x2, a_major=MN_MAJOR, b_major=MN_MAJOR, out=dst, accumulate=acc))
        if fuse_db1:
            def _finalize(dst, acc):
                _lib.check(_lib.cuda().mg_colsum_finalize(cs_part.data_ptr(), cs_rows, N1, dst.data_ptr(),
                                                          1 if acc else 0, _stream()), "mg_colsum_finalize")
                _count(1)

            db1 = _write_or_return_grad(b1, _finalize)
        else:
            db1 = _write_or_return_grad(b1, lambda dst, acc: colsum(dpre, dst, accumulate=acc))
        return dx, dw1, db1, dw2, db2, (dy if ctx.has_res else None)


def mlp(x, w1, b1, w2, b2, residual=None):
    return _MLP.apply(x, w1, b1, w2, b2, residual)


class _CrossEntropy(torch.autograd.Function):
    """mean NLL; the logits buffer is overwritten with its own gradient.

    ``count_valid=True`` is torch's ``ignore_index`` semantics (mean over NON-ignored rows): the divisor is counted
    on the device and handed to the kernel as a device scalar, so no host sync is needed. ``False`` divides by the
    total row count (identical when no target is ignored, and one reduction cheaper)."""

    @staticmethod
    def forward(ctx, logits, targets, v_real, ignore_index, unit_grad, count_valid):
        l2 = logits.reshape(-1, logits.shape[-1])
        assert l2.stride(1) == 1
        t = targets.reshape(-1)
        rows, v_pad = l2.shape
        losses = torch.empty(rows, dtype=torch.float32, device=logits.device)
        inv = None
        if count_valid:
            valid = ((t != ignore_index) & (t >= 0) & (t < v_real)).sum()
            inv = valid.clamp_(min=1).to(torch.float32).reciprocal_()
        rc = _lib.cuda().mg_ce_fwd_bwd(l2.data_ptr(), t.data_ptr(), losses.data_ptr(), rows, v_real,
                                       v_pad, l2.stride(0), 1.0 if count_valid else 1.0 / rows, _ptr(inv),
                                       ignore_index, _stream())
        _lib.check(rc, "mg_ce_fwd_bwd")
        _count(1)
        ctx.save_for_backward(l2)
        ctx.lshape = logits.shape
        ctx.unit = unit_grad
        return losses.sum() * inv if count_valid else losses.sum() / rows

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        # unit_grad=True (the training-step case: loss.backward()) skips a 2x|logits| scaling pass
        g = dlogits.view(ctx.lshape)
        if not ctx.unit:
            g = g * dloss.to(g.dtype)
        return g, None, None, None, None, None


def cross_entropy(logits, targets, v_real=None, ignore_index=-100, unit_grad=True, reduction="mean"):
    """Fused softmax-CE. NOTE: consumes ``logits`` (overwritten by d loss/d logits).

    ``reduction="mean"`` averages over the rows whose target is not ``ignore_index`` (torch semantics);
    ``"mean_all"`` divides by the total row count — the same number when nothing is ignored (causal-LM training),
    without the device-side count. ``unit_grad=True`` assumes the loss is the root of backward."""
    if reduction not in ("mean", "mean_all"):
        raise ValueError(f"reduction must be 'mean' or 'mean_all', got {reduction!r}")
    return _CrossEntropy.apply(logits, targets, v_real or logits.shape[-1], ignore_index, unit_grad,
                               reduction == "mean")


class _Embedding(torch.autograd.Function):
    """out[b, t] = wte[idx[b, t]] (+ wpe[t]); ``wpe`` may be None (rotary-position models)."""

    @staticmethod
    def forward(ctx, idx, wte, wpe):
        B, T = idx.shape
        C = wte.shape[1]
        out = torch.empty((B, T, C), dtype=torch.bfloat16, device=wte.device)
        rc = _lib.cuda().mg_embed_fwd(idx.data_ptr(), wte.data_ptr(), _ptr(wpe), out.data_ptr(),
                                      B * T, T, C, _stream())
        _lib.check(rc, "mg_embed_fwd")
        _count(1)
        ctx.save_for_backward(idx)
        ctx.refs = (wte, wpe)
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        wte, wpe = ctx.refs
        B, T = idx.shape
        C = wte.shape[1]
        dout = dout.contiguous()
        use_main = getattr(wte, "main_grad", None) is not None
        dwpe, acc_wpe = None, False
        if use_main:
            dwte = wte.main_grad
            if not getattr(wte, "_mg_written", False):
                dwte.zero_()
            wte._mg_written = True
            if wpe is not None:
                dwpe = wpe.main_grad
                acc_wpe = bool(getattr(wpe, "_mg_written", False))
                if not acc_wpe and T < wpe.shape[0]:
                    dwpe[T:].zero_()
                wpe._mg_written = True
        else:
            dwte = torch.zeros_like(wte)
            if wpe is not None:
                dwpe = torch.zeros_like(wpe) if T < wpe.shape[0] else torch.empty_like(wpe)
        rc = _lib.cuda().mg_embed_bwd(idx.data_ptr(), dout.data_ptr(), dwte.data_ptr(), _ptr(dwpe),
                                      B, T, C, 1 if acc_wpe else 0, _stream())
        _lib.check(rc, "mg_embed_bwd")
        _count(2 if dwpe is not None else 1)
        if use_main:
            return None, None, None
        return None, dwte, dwpe


def embedding(idx, wte, wpe=None):
    return _Embedding.apply(idx, wte, wpe)
