"""ZeRO optimizer patching (reference: maggy/core/patching/optim.py:28-117 and
torch_dist_executor.py:408-422, which swap ``torch.optim.*`` for ZeRO subclasses process-wide).

Here the swap is *scoped*: ``patched_optimizers(zero_lvl)`` is a context manager active only while
train_fn runs. Inside it ``torch.optim.Adam`` / ``AdamW`` constructed on the parameters of a fused
``DataParallel`` module return that module's :class:`~maggy_b200.parallel.FusedAdamW` (reduce-scatter +
Adam + all-gather in one NVLink kernel); every other optimizer class / module gets
``ZeroRedundancyOptimizer`` (ZeRO-1) exactly like the reference.
"""
from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist

_ZERO_CLASSES = ["Adadelta", "Adagrad", "Adam", "AdamW", "SparseAdam", "Adamax", "ASGD", "LBFGS", "RMSprop",
                 "Rprop", "SGD"]


def _find_fused_owner(params):
    """The DataParallel wrapper that owns these parameters (they carry a ``main_grad`` arena view)."""
    params = list(params)
    flat = None
    for p in params:
        if isinstance(p, dict):
            return None, params
        owner = getattr(p, "_maggy_flat", None)
        if owner is None:
            return None, params
        flat = owner
    return flat, params


class MaggyZeroOptimizer:
    """Marker base of every ``MaggyZero<Name>`` class (reference: patching/optim.py:28-45 subclasses torch's
    ZeroRedundancyOptimizer; here construction dispatches to the fused arena optimizer, ZeRO-1 or the plain optimizer,
    see :func:`make_zero_optimizer`). ``optim_cls`` is bound by the generated subclasses."""

    optim_cls = None


def make_zero_optimizer(optim_cls, zero_lvl: int = 1):
    """``MaggyZero<Name>``: same constructor as ``optim_cls``. Parameters that live in a fused ``DataParallel`` arena
    get the fused (and sharded) path at every ZeRO level — there is no DDP underneath that would reduce their
    gradients otherwise; other parameters get ``ZeroRedundancyOptimizer`` for ``zero_lvl`` 1-2 like the reference and
    the plain optimizer for level 0."""

    _bound = optim_cls

    class _MaggyZero(MaggyZeroOptimizer):
        optim_cls = _bound

        def __new__(cls, params, *args, **kwargs):
            flat, params = _find_fused_owner(params)
            # (by name: inside patched_optimizers() torch.optim.Adam IS the patched class, not the saved original)
            if flat is not None and optim_cls.__name__ in ("Adam", "AdamW"):
                from ...parallel.arena import FusedAdamW

                names = ("lr", "betas", "eps", "weight_decay")
                kw = dict(zip(names, args))
                kw.update({k: v for k, v in kwargs.items() if k in names})
                if optim_cls.__name__ == "Adam":
                    kw.setdefault("weight_decay", 0.0)
                return FusedAdamW(flat, **kw)
            if flat is not None:
                # the parameters live in a fused DataParallel arena but the optimizer is not Adam(W): run the user's
                # optimizer class unchanged and average the whole flat gradient arena with ONE collective before
                # every step (there is no DDP underneath to do it bucket by bucket)
                opt = optim_cls(params, *args, **kwargs)

                def _reduce(optimizer, a, k, _flat=flat):
                    _flat.adopt_autograd_grads()
                    if _flat.world_size > 1:
                        dist.all_reduce(_flat.grads, op=dist.ReduceOp.AVG, group=_flat.group)

                opt.register_step_pre_hook(_reduce)
                _orig_zero = opt.zero_grad

                def _zero(set_to_none: bool = True, _flat=flat):
                    _flat.start_step()     # gradients live in the arena: zero in place, keep p.grad pointing at it

                opt.zero_grad = _zero
                return opt
            if zero_lvl in (1, 2) and dist.is_initialized() and dist.get_world_size() > 1:
                from torch.distributed.optim import ZeroRedundancyOptimizer

                names = list(optim_cls.__init__.__code__.co_varnames[2:optim_cls.__init__.__code__.co_argcount])
                kw = dict(zip(names, args))
                kw.update(kwargs)
                return ZeroRedundancyOptimizer(params, optimizer_class=optim_cls, **kw)
            return optim_cls(params, *args, **kwargs)

    _MaggyZero.__name__ = _MaggyZero.__qualname__ = f"MaggyZero{optim_cls.__name__}"
    return _MaggyZero


MaggyZeroAdadelta, MaggyZeroAdagrad, MaggyZeroAdam, MaggyZeroAdamW, MaggyZeroSparseAdam, MaggyZeroAdamax, \
    MaggyZeroASGD, MaggyZeroLBFGS, MaggyZeroRMSprop, MaggyZeroRprop, MaggyZeroSGD = [
        make_zero_optimizer(getattr(torch.optim, n)) for n in _ZERO_CLASSES]


@contextlib.contextmanager
def patched_optimizers(zero_lvl: int):
    """Swap ``torch.optim.<X>`` for the maggy variants while the block runs (scoped, not process-wide)."""
    saved = {n: getattr(torch.optim, n) for n in _ZERO_CLASSES}
    try:
        for n in _ZERO_CLASSES:
            setattr(torch.optim, n, make_zero_optimizer(saved[n], zero_lvl))
        yield
    finally:
        for n, cls in saved.items():
            setattr(torch.optim, n, cls)
