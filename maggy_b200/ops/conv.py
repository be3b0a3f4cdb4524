"""conv2d on the tcgen05 kernels (K7): NHWC bf16 activations, KRSC bf16 filters.

Replaces the cuDNN convolutions the reference reaches through the user's ``train_fn`` in the ASHA ResNet-50
config (reference: maggy/core/executors/trial_executor.py:126-140 runs the oblivious
function, BASELINE.json names the ResNet-50 config).

Three routes, all on hand-written kernels:

* **implicit GEMM** (``csrc/ops/conv_tcgen05.cu``) for stride-1 convolutions with ``C % 64 == 0`` and
  ``Kout % 64 == 0``: forward and dgrad are the same kernel (dgrad runs it on ``dY`` with the flipped, transposed
  filter), wgrad is a pixel-reduction GEMM with both operands loaded as shifted 4-D TMA boxes;
* **1x1, stride 1**: the activation tensor *is* the GEMM operand — ``ops.functional.linear`` on ``[N*H*W, C]``;
* **explicit** (strided or thin-channel convolutions, e.g. the 7x7/2 stem): ``im2col`` gather kernel -> tcgen05
  GEMM, backward ``col2im`` gather-sum; the column matrix is kept for the wgrad GEMM (``MAGGY_B200_CONV_SAVE_COLS=0``
  recomputes it instead). Thin-channel inputs use a shared-memory staged gather that writes whole 16-byte vectors.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

from . import _lib
from . import functional as F

_P = _lib.c_void_p
_I = _lib.c_int
_SAVE_COLS = __import__("os").environ.get("MAGGY_B200_CONV_SAVE_COLS", "1") != "0"


def _register():
    _lib.register_cuda_symbol("mg_conv2d_fprop", [_P, _P, _P] + [_I] * 11 + [_P])
    _lib.register_cuda_symbol("mg_conv2d_wgrad", [_P, _P, _P] + [_I] * 11 + [_P])
    _lib.register_cuda_symbol("mg_im2col_nhwc", [_P, _P] + [_I] * 10 + [_P])
    _lib.register_cuda_symbol("mg_col2im_nhwc", [_P, _P] + [_I] * 10 + [_P])


_registered = False


def _cuda():
    global _registered
    lib = _lib.cuda()
    if not _registered:
        _register()
        _registered = True
    return lib


def _pair(v) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def implicit_ok(C: int, Kout: int, stride: int) -> bool:
    return stride == 1 and C % 64 == 0 and Kout % 64 == 0


# ------------------------------------------------------------------------------------------------
# raw launchers (no autograd)
# ------------------------------------------------------------------------------------------------
def conv2d_fprop(x: torch.Tensor, w: torch.Tensor, padding=0, relu: bool = False, sm_limit: int = 0):
    """Implicit-GEMM forward. ``x``: [N,H,W,C] bf16 contiguous; ``w``: [Kout,R,S,C] bf16 contiguous; stride 1."""
    F._need_cuda(x, w)
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.is_contiguous() and w.is_contiguous()
    N, H, W, C = x.shape
    Kout, R, S, Cw = w.shape
    assert C == Cw, f"channel mismatch {C} vs {Cw}"
    ph, pw = _pair(padding)
    P, Q = H + 2 * ph - R + 1, W + 2 * pw - S + 1
    y = torch.empty((N, P, Q, Kout), dtype=torch.bfloat16, device=x.device)
    rc = _cuda().mg_conv2d_fprop(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, H, W, C, Kout, R, S, ph, pw,
                                 1 if relu else 0, sm_limit, F._stream())
    _lib.check(rc, f"mg_conv2d_fprop(N={N},H={H},W={W},C={C},K={Kout},R={R},S={S})")
    F._count(1)
    return y


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, R: int, S: int, padding=0, splits: int = 0, sm_limit: int = 0):
    """Implicit-GEMM weight gradient, fp32 ``[Kout,R,S,C]``. ``x``: [N,H,W,C]; ``dy``: [N,P,Q,Kout]; stride 1."""
    F._need_cuda(x, dy)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
    N, H, W, C = x.shape
    Kout = dy.shape[3]
    ph, pw = _pair(padding)
    assert dy.shape[1] == H + 2 * ph - R + 1 and dy.shape[2] == W + 2 * pw - S + 1
    dw = torch.empty((Kout, R, S, C), dtype=torch.float32, device=x.device)
    rc = _cuda().mg_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W, C, Kout, R, S, ph, pw,
                                 splits, sm_limit, F._stream())
    _lib.check(rc, f"mg_conv2d_wgrad(N={N},H={H},W={W},C={C},K={Kout},R={R},S={S})")
    F._count(2)
    return dw


def im2col(x: torch.Tensor, R: int, S: int, stride: int, padding) -> torch.Tensor:
    """``[N*P*Q, Kp]`` column matrix (``Kp = ceil8(R*S*C)``, zero tail) of the NHWC tensor ``x``."""
    F._need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    N, H, W, C = x.shape
    ph, pw = _pair(padding)
    P, Q = (H + 2 * ph - R) // stride + 1, (W + 2 * pw - S) // stride + 1
    Kp = (R * S * C + 7) // 8 * 8
    cols = torch.empty((N * P * Q, Kp), dtype=torch.bfloat16, device=x.device)
    rc = _cuda().mg_im2col_nhwc(x.data_ptr(), cols.data_ptr(), N, H, W, C, R, S, stride, ph, pw, Kp, F._stream())
    _lib.check(rc, "mg_im2col_nhwc")
    F._count(1)
    return cols


def col2im(dcols: torch.Tensor, xshape, R: int, S: int, stride: int, padding) -> torch.Tensor:
    F._need_cuda(dcols)
    N, H, W, C = xshape
    ph, pw = _pair(padding)
    dx = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=dcols.device)
    rc = _cuda().mg_col2im_nhwc(dcols.data_ptr(), dx.data_ptr(), N, H, W, C, R, S, stride, ph, pw, dcols.shape[1],
                                F._stream())
    _lib.check(rc, "mg_col2im_nhwc")
    F._count(1)
    return dx


# ------------------------------------------------------------------------------------------------
# autograd
# ------------------------------------------------------------------------------------------------
def _padded_filter(w: torch.Tensor, Kp: int) -> torch.Tensor:
    Kout = w.shape[0]
    w2 = w.reshape(Kout, -1)
    if w2.shape[1] == Kp:
        return w2
    out = torch.zeros((Kout, Kp), dtype=w.dtype, device=w.device)
    out[:, : w2.shape[1]] = w2
    return out


class _Conv2d(torch.autograd.Function):
    """x: [N,H,W,C] bf16, weight2d: [Kout, R*S*C] bf16 (KRSC flattened) -> [N,P,Q,Kout] bf16."""

    @staticmethod
    def forward(ctx, x, weight2d, R, S, stride, padding, with_skip=False):
        if not x.is_contiguous():
            x = x.contiguous()
        ctx.with_skip = with_skip
        y = _Conv2d._forward(ctx, x, weight2d, R, S, stride, padding)
        # with_skip: also hand back x itself (an alias, no kernel) for the skip connection around this convolution, so
        # the skip gradient arrives in THIS backward and is added in the dgrad GEMM epilogue (no separate add kernel)
        return (y, x.view_as(x)) if with_skip else y

    @staticmethod
    def _forward(ctx, x, weight2d, R, S, stride, padding):
        Kout = weight2d.shape[0]
        C = weight2d.shape[1] // (R * S)
        assert x.shape[3] == C and weight2d.is_contiguous()
        weight = weight2d.view(Kout, R, S, C)
        ctx.R, ctx.S = R, S
        ctx.cols = None
        ctx.stride, ctx.padding = stride, _pair(padding)
        ctx.pointwise = R == 1 and S == 1 and stride == 1 and ctx.padding == (0, 0)
        ctx.implicit = implicit_ok(C, Kout, stride)
        ctx.weight_ref = weight2d
        ctx.save_for_backward(x, weight2d)
        if ctx.pointwise:
            # the activation tensor is the GEMM operand as it stands
            N, H, W, _ = x.shape
            return F.gemm(x.view(-1, C), weight.view(Kout, C)).view(N, H, W, Kout)
        if ctx.implicit:
            return conv2d_fprop(x, weight, ctx.padding)
        cols = im2col(x, R, S, stride, ctx.padding)
        N, H, W, _ = x.shape
        P = (H + 2 * ctx.padding[0] - R) // stride + 1
        Q = (W + 2 * ctx.padding[1] - S) // stride + 1
        y = F.gemm(cols, _padded_filter(weight, cols.shape[1]))
        # the column matrix is the wgrad operand: keep it (default; 244 MB for the ResNet-50 stem at batch 64) instead of
        # gathering it a second time in backward. MAGGY_B200_CONV_SAVE_COLS=0 trades the memory back for the recompute.
        ctx.cols = cols if _SAVE_COLS and ctx.needs_input_grad[1] else None
        return y.view(N, P, Q, Kout)

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x, weight2d = ctx.saved_tensors
        R, S = ctx.R, ctx.S
        Kout, C = weight2d.shape[0], weight2d.shape[1] // (R * S)
        weight = weight2d.view(Kout, R, S, C)
        if dy is None:  # only the skip connection carried a gradient
            return dskip, None, None, None, None, None, None
        if not dy.is_contiguous():
            dy = dy.contiguous()
        if dskip is not None and not dskip.is_contiguous():
            dskip = dskip.contiguous()
        dx = None
        if ctx.pointwise:
            dy2, x2 = dy.view(-1, Kout), x.view(-1, C)
            if ctx.needs_input_grad[0]:
                if dskip is not None:   # dx = dY W + dskip in one epilogue
                    dx = F.gemm(dy2, weight.view(Kout, C), b_major=F.MN_MAJOR, aux=dskip.view(-1, C),
                                epi=F.EPI_BIAS_RESIDUAL).view(x.shape)
                else:
                    dx = F.gemm(dy2, weight.view(Kout, C), b_major=F.MN_MAJOR).view(x.shape)
            dw = F._write_or_return_grad(
                ctx.weight_ref,
                lambda dst, acc: F.gemm(dy2, x2, a_major=F.MN_MAJOR, b_major=F.MN_MAJOR, out=dst, accumulate=acc))
            return dx, dw, None, None, None, None, None
        if ctx.implicit:
            ph, pw = ctx.padding
            if ctx.needs_input_grad[0]:
                # dX = conv(dY, flip(W)^T) with padding R-1-pad: the forward kernel again
                wt = weight.flip(1, 2).permute(3, 1, 2, 0).contiguous()
                dx = conv2d_fprop(dy, wt, (R - 1 - ph, S - 1 - pw))
            dw32 = conv2d_wgrad(x, dy, R, S, ctx.padding).view(Kout, R * S * C)

            def into(dst, acc):
                if acc:
                    dst.add_(dw32)
                else:
                    dst.copy_(dw32)
        else:
            cols = ctx.cols if ctx.cols is not None else im2col(x, R, S, ctx.stride, ctx.padding)
            ctx.cols = None
            dy2 = dy.view(-1, Kout)
            if ctx.needs_input_grad[0]:
                dcols = F.gemm(dy2, _padded_filter(weight, cols.shape[1]), b_major=F.MN_MAJOR)
                dx = col2im(dcols, x.shape, R, S, ctx.stride, ctx.padding)
                del dcols
            dw2 = F.gemm(dy2, cols, a_major=F.MN_MAJOR, b_major=F.MN_MAJOR, out_dtype=torch.float32)
            dw32 = dw2[:, : R * S * C]

            def into(dst, acc):
                if acc:
                    dst.add_(dw32)
                else:
                    dst.copy_(dw32)
        dw = F._write_or_return_grad(ctx.weight_ref, into)
        if dskip is not None:
            dx = dskip if dx is None else dx + dskip
        return dx, dw, None, None, None, None, None


def conv2d(x: torch.Tensor, weight: torch.Tensor, stride: int = 1, padding=0, kernel_size=None, with_skip=False):
    """NHWC convolution: ``x`` [N,H,W,C] bf16, ``weight`` [Kout,R,S,C] (or [Kout, R*S*C] with ``kernel_size``) bf16
    -> [N,P,Q,Kout] bf16. ``with_skip=True`` returns ``(y, x_skip)``: feed ``x_skip`` (an alias of ``x``) into the
    skip connection around the convolution and its gradient is added inside the dgrad epilogue."""
    if weight.dim() == 4:
        Kout, R, S, C = weight.shape
        assert weight.is_contiguous(), "KRSC filter must be contiguous"
        weight = weight.view(Kout, R * S * C)
    else:
        R, S = _pair(kernel_size)
    return _Conv2d.apply(x, weight, R, S, stride, padding, with_skip)


class Conv2d(torch.nn.Module):
    """Drop-in for ``nn.Conv2d(bias=False)`` on logical-NCHW tensors in ``channels_last`` memory format.

    The filter parameter is stored as the 2-D GEMM operand ``[out, kh*kw*in]`` (KRSC flattened — the layout the
    kernels read, and immune to ``module.to(memory_format=channels_last)``); :meth:`load_torch_weight` /
    :meth:`torch_weight` convert from / to torch's ``[out, in, kh, kw]``. On CPU tensors (unit tests without a GPU)
    the module evaluates ``torch.nn.functional.conv2d`` — on CUDA tensors only the native kernels run.
    """

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride: int = 1, padding=0,
                 bias: bool = False, dtype=torch.bfloat16):
        super().__init__()
        assert not bias, "fold the bias into the following normalisation layer"
        kh, kw = _pair(kernel_size)
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, (kh, kw)
        self.stride, self.padding = int(stride), _pair(padding)
        self.weight = torch.nn.Parameter(torch.empty(out_channels, kh * kw * in_channels, dtype=dtype))
        self.reset_parameters()

    def reset_parameters(self):
        kh, kw = self.kernel_size
        std = math.sqrt(2.0 / (self.out_channels * kh * kw))  # kaiming_normal_, fan_out, relu
        with torch.no_grad():
            self.weight.copy_(torch.randn(self.weight.shape, dtype=torch.float32) * std)

    def load_torch_weight(self, w_oihw: torch.Tensor):
        with torch.no_grad():
            self.weight.copy_(w_oihw.permute(0, 2, 3, 1).reshape(self.out_channels, -1))

    def torch_weight(self) -> torch.Tensor:
        kh, kw = self.kernel_size
        return self.weight.view(self.out_channels, kh, kw, self.in_channels).permute(0, 3, 1, 2)

    def forward(self, x: torch.Tensor, with_skip: bool = False):
        if not x.is_cuda:
            y = torch.nn.functional.conv2d(x, self.torch_weight().to(x.dtype), None, self.stride, self.padding)
            return (y, x) if with_skip else y
        xn = x.permute(0, 2, 3, 1)  # NHWC view of a channels_last tensor
        if xn.dtype != torch.bfloat16:
            xn = xn.to(torch.bfloat16)
        if with_skip:
            y, xs = conv2d(xn, self.weight, self.stride, self.padding, kernel_size=self.kernel_size, with_skip=True)
            return y.permute(0, 3, 1, 2), xs.permute(0, 3, 1, 2)
        y = conv2d(xn, self.weight, self.stride, self.padding, kernel_size=self.kernel_size)
        return y.permute(0, 3, 1, 2)  # logical NCHW, channels_last memory

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, layout=KRSC/NHWC")
