"""Training BatchNorm on NHWC bf16 activations, fused with ReLU and the residual add (K7b, csrc/ops/batchnorm.cu).

The ResNet-50 ASHA trial body (BASELINE.json configs 1-2; the reference just runs the user's ``train_fn``:
maggy/core/executors/trial_executor.py:126-140) spends more than half of its step in ATen's channels-last BatchNorm
kernels; :class:`BatchNorm2d` replaces them — and the separate ReLU / add kernels around them — with three streaming
kernels per direction. Drop-in: subclasses ``torch.nn.BatchNorm2d`` (same parameters, buffers, ``state_dict``), falls
back to ATen for anything the kernels do not cover (CPU tensors, fp32 activations, NCHW-contiguous memory,
``momentum=None``).
"""
from __future__ import annotations

import torch

from . import _lib
from . import functional as F

_P, _I, _LL, _FL = _lib.c_void_p, _lib.c_int, _lib.c_ll, _lib.c_float
_registered = False


def _cuda():
    global _registered
    lib = _lib.cuda()
    if not _registered:
        _lib.register_cuda_symbol("mg_bn_partial_floats", [_I], restype=_LL)
        _lib.register_cuda_symbol("mg_bn_fwd_stats", [_P, _LL, _I, _P, _P, _I, _P, _P, _FL, _FL, _P, _P, _P, _P, _P])
        _lib.register_cuda_symbol("mg_bn_eval_coef", [_P, _P, _P, _P, _I, _FL, _I, _P, _P])
        _lib.register_cuda_symbol("mg_bn_apply", [_P, _P, _P, _P, _LL, _I, _I, _P])
        _lib.register_cuda_symbol("mg_bn_bwd", [_P, _P, _P, _P, _P, _I, _P, _P, _I, _P, _P, _LL, _I, _I, _P, _P, _P])
        _registered = True
    return lib


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _param_flag(p) -> int:
    return 1 if p is not None and p.dtype == torch.bfloat16 else 0


def supported(x: torch.Tensor, weight, running_mean, momentum) -> bool:
    """True when the native kernels cover this call (``x`` channels-last-contiguous ``[..., C]`` bf16 on CUDA)."""
    C = x.shape[-1]
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and C % 8 == 0 and C <= 2048
            and momentum is not None
            and (weight is None or weight.dtype in (torch.float32, torch.bfloat16))
            and (running_mean is None or running_mean.dtype == torch.float32))


class _BatchNormAct(torch.autograd.Function):
    """y = relu?(batch_norm(x) (+ residual)) over the last dimension of a contiguous ``[..., C]`` bf16 tensor."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, num_batches, training, momentum, eps, relu):
        lib = _cuda()
        C = x.shape[-1]
        M = x.numel() // C
        dev = x.device
        if residual is not None:
            assert residual.shape == x.shape and residual.dtype == x.dtype
            if not residual.is_contiguous():
                residual = residual.contiguous()
        y = torch.empty_like(x)
        coef = torch.empty((2, C), dtype=torch.float32, device=dev)
        stats = None
        flag = _param_flag(weight if weight is not None else bias)
        stream = F._stream()
        if training:
            stats = torch.empty((2, C), dtype=torch.float32, device=dev)
            partial = F._scratch_f32("bn_partial", int(lib.mg_bn_partial_floats(C)), dev)
            rc = lib.mg_bn_fwd_stats(x.data_ptr(), M, C, _ptr(weight), _ptr(bias), flag, _ptr(running_mean),
                                     _ptr(running_var), float(momentum), float(eps), stats.data_ptr(), coef.data_ptr(),
                                     _ptr(num_batches), partial.data_ptr(), stream)
            _lib.check(rc, "mg_bn_fwd_stats")
            F._count(2)
        else:
            rc = lib.mg_bn_eval_coef(running_mean.data_ptr(), running_var.data_ptr(), _ptr(weight), _ptr(bias), flag,
                                     float(eps), C, coef.data_ptr(), stream)
            _lib.check(rc, "mg_bn_eval_coef")
            F._count(1)
        rc = lib.mg_bn_apply(x.data_ptr(), _ptr(residual), y.data_ptr(), coef.data_ptr(), M, C, 1 if relu else 0, stream)
        _lib.check(rc, "mg_bn_apply")
        F._count(1)
        ctx.training, ctx.relu, ctx.has_res = training, relu, residual is not None
        ctx.w_ref, ctx.b_ref = weight, bias
        ctx.eps = eps
        if training:
            ctx.save_for_backward(x, y if relu else None, stats)
        else:
            ctx.save_for_backward(x, y if relu else None, running_mean, running_var)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _cuda()
        saved = ctx.saved_tensors
        x, y = saved[0], saved[1]
        C = x.shape[-1]
        M = x.numel() // C
        dev = x.device
        if not dy.is_contiguous():
            dy = dy.contiguous()
        w, b = ctx.w_ref, ctx.b_ref
        if ctx.training:
            stats = saved[2]
        else:  # frozen statistics: dx = gamma / std * dz — expressed through the same kernels with dgamma/dbeta terms off
            rm, rv = saved[2], saved[3]
            stats = torch.stack([rm, torch.rsqrt(rv + ctx.eps)]).contiguous()
        use_main = w is not None and getattr(w, "main_grad", None) is not None
        dgamma = dbeta = None
        acc = False
        if w is not None and ctx.needs_input_grad[1]:
            if use_main:
                acc = bool(getattr(w, "_mg_written", False))
                dgamma, dbeta = w.main_grad, (b.main_grad if b is not None else None)
                w._mg_written = True
                if b is not None:
                    b._mg_written = True
            else:
                dgamma = torch.empty_like(w)
                dbeta = torch.empty_like(b) if b is not None else None
        dx = torch.empty_like(x)
        dz = torch.empty_like(x) if (ctx.has_res and ctx.relu) else None
        coef = torch.empty((3, C), dtype=torch.float32, device=dev)
        partial = F._scratch_f32("bn_partial", int(lib.mg_bn_partial_floats(C)), dev)
        rc = lib.mg_bn_bwd(dy.data_ptr(), _ptr(y), x.data_ptr(), stats.data_ptr(), _ptr(w), _param_flag(w),
                           _ptr(dgamma), _ptr(dbeta), 1 if acc else 0, dx.data_ptr(), _ptr(dz), M, C,
                           1 if ctx.relu else 0, coef.data_ptr(), partial.data_ptr(), F._stream())
        _lib.check(rc, "mg_bn_bwd")
        F._count(3)
        if not ctx.training:
            # statistics are constants: the mean / variance terms of dx vanish -> dx = a * dz (recomputed cheaply)
            dzt = dz if dz is not None else (dy * (y > 0) if ctx.relu else dy)
            dx = (dzt.float() * coef[0]).to(x.dtype)
        dres = (dz if ctx.relu else dy) if ctx.has_res else None
        if use_main:
            dgamma = dbeta = None
        return dx, dgamma, dbeta, dres, None, None, None, None, None, None, None


def batch_norm_act(x, weight, bias, running_mean=None, running_var=None, num_batches=None, training=True,
                   momentum=0.1, eps=1e-5, relu=False, residual=None):
    """``relu?(batch_norm(x) + residual)`` over the last (channel) dimension of a contiguous bf16 tensor."""
    F._need_cuda(x, residual)
    use_batch_stats = training or running_mean is None
    return _BatchNormAct.apply(x, weight, bias, residual, running_mean, running_var, num_batches if training else None,
                               use_batch_stats, momentum, eps, relu)


class BatchNorm2d(torch.nn.BatchNorm2d):
    """``nn.BatchNorm2d`` for logical-NCHW tensors in ``channels_last`` memory, optionally fused with the ReLU that
    follows it (``relu=True``) and with a residual add before that ReLU (``forward(x, residual)``)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, relu=False,
                 device=None, dtype=None):
        super().__init__(num_features, eps, momentum, affine, track_running_stats, device=device, dtype=dtype)
        self.relu = relu

    def forward(self, x, residual=None):
        native = (x.dim() == 4 and x.is_cuda and x.is_contiguous(memory_format=torch.channels_last)
                  and (residual is None or (residual.dtype == x.dtype and residual.shape == x.shape
                                            and residual.is_contiguous(memory_format=torch.channels_last))))
        if native:
            xn = x.permute(0, 2, 3, 1)
            native = supported(xn, self.weight, self.running_mean, self.momentum)
        if not native:
            out = super().forward(x)
            if residual is not None:
                out = out + residual
            return torch.relu(out) if self.relu else out
        rn = residual.permute(0, 2, 3, 1) if residual is not None else None
        y = batch_norm_act(xn, self.weight, self.bias, self.running_mean, self.running_var, self.num_batches_tracked,
                           self.training, self.momentum, self.eps, self.relu, rn)
        return y.permute(0, 3, 1, 2)

    def extra_repr(self):
        return super().extra_repr() + f", relu={self.relu}, layout=NHWC"
