"""Max pooling on NHWC bf16 activations (K7c, csrc/ops/pool.cu): the pool after the ResNet stem in the ASHA trial body
(BASELINE.json configs 1-2; the reference runs the user's train_fn as is: maggy/core/executors/trial_executor.py:126-140).
:class:`MaxPool2d` is a drop-in for ``nn.MaxPool2d`` on channels-last tensors and falls back to ATen elsewhere."""
from __future__ import annotations

import torch

from . import _lib
from . import functional as F

_P, _I = _lib.c_void_p, _lib.c_int
_registered = False


def _cuda():
    global _registered
    lib = _lib.cuda()
    if not _registered:
        _lib.register_cuda_symbol("mg_maxpool_fwd", [_P, _P, _P] + [_I] * 7 + [_P])
        _lib.register_cuda_symbol("mg_maxpool_bwd", [_P, _P, _P] + [_I] * 7 + [_P])
        _registered = True
    return lib


class _MaxPoolNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        N, H, W, C = x.shape
        P, Q = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty((N, P, Q, C), dtype=x.dtype, device=x.device)
        idx = torch.empty((N, P, Q, C), dtype=torch.uint8, device=x.device)
        rc = _cuda().mg_maxpool_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr(), N, H, W, C, k, stride, pad, F._stream())
        _lib.check(rc, "mg_maxpool_fwd")
        F._count(1)
        ctx.save_for_backward(idx)
        ctx.geom = (N, H, W, C, k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, H, W, C, k, stride, pad = ctx.geom
        if not dy.is_contiguous():
            dy = dy.contiguous()
        dx = torch.empty((N, H, W, C), dtype=dy.dtype, device=dy.device)
        rc = _cuda().mg_maxpool_bwd(dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), N, H, W, C, k, stride, pad, F._stream())
        _lib.check(rc, "mg_maxpool_bwd")
        F._count(1)
        return dx, None, None, None


def max_pool_nhwc(x: torch.Tensor, kernel_size: int, stride: int, padding: int = 0) -> torch.Tensor:
    """``x``: contiguous ``[N, H, W, C]`` bf16 on CUDA, ``C % 8 == 0``."""
    F._need_cuda(x)
    return _MaxPoolNHWC.apply(x, int(kernel_size), int(stride), int(padding))


class MaxPool2d(torch.nn.MaxPool2d):
    def forward(self, x):
        k, s, p = self.kernel_size, self.stride, self.padding
        native = (x.dim() == 4 and x.is_cuda and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0
                  and x.is_contiguous(memory_format=torch.channels_last) and isinstance(k, int) and isinstance(s, int)
                  and isinstance(p, int) and self.dilation == 1 and not self.ceil_mode and not self.return_indices
                  and k * k <= 255 and 2 * p <= k)
        if not native:
            return super().forward(x)
        return max_pool_nhwc(x.permute(0, 2, 3, 1), k, s, p).permute(0, 3, 1, 2)
