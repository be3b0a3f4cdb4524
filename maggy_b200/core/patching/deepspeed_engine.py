"""DeepSpeed-engine semantics on the maggy_b200 data plane (reference: ``MaggyDeepSpeedModuleWrapper(DeepSpeedEngine)``,
maggy/core/patching/modules.py:100-139, and the config sanitising of torch_dist_executor.py:389-405).

DeepSpeed itself is not used. The class the train_fn receives behaves like the engine the reference hands out:

    engine = module(*ctor_args)        # builds the user's module, wraps it, creates the optimizer from the config
    loss = engine(batch)               # forward
    engine.backward(loss)              # scales the loss (fp16 loss scaling) and runs autograd
    engine.step()                      # at an accumulation boundary: overflow check -> optimizer step -> zero grads

Honoured config keys: ``zero_optimization.stage`` (0-2: fused reduce-scatter + sharded Adam, K1; 3: sharded parameters
with the all-gather fused into the GEMM, K3), ``optimizer`` (Adam / AdamW only — the reference forces Adam too —
``params``: lr, betas, eps, weight_decay), ``fp16`` (enabled, loss_scale [0 = dynamic], initial_scale_power,
loss_scale_window, hysteresis, min_loss_scale), ``bf16.enabled``, ``gradient_accumulation_steps``,
``gradient_clipping``, ``train_batch_size`` / ``train_micro_batch_size_per_gpu`` (consistency check only).
fp16 configs compute in bf16 on B200 (the tensor cores' native 16-bit training type); the loss-scaling state machine is
kept so that engine.loss_scale / skipped_steps behave as DeepSpeed users expect.
"""
from __future__ import annotations

from typing import Optional, Type

import torch
import torch.distributed as dist


class DynamicLossScaler:
    """DeepSpeed's dynamic loss scaler: halve on overflow (after ``hysteresis`` consecutive overflows), double after
    ``scale_window`` clean steps, never below ``min_scale``."""

    def __init__(self, init_scale=2.0 ** 16, scale_window=1000, hysteresis=2, min_scale=1.0, dynamic=True):
        self.cur_scale = float(init_scale)
        self.scale_window, self.min_scale, self.dynamic = int(scale_window), float(min_scale), dynamic
        self.delayed_shift = int(hysteresis)
        self.cur_hysteresis = int(hysteresis)
        self.last_overflow_iter = -1
        self.cur_iter = 0

    def update(self, overflow: bool) -> None:
        if not self.dynamic:
            self.cur_iter += 1
            return
        if overflow:
            if self.delayed_shift == 1 or self.cur_hysteresis == 1:
                self.cur_scale = max(self.cur_scale / 2.0, self.min_scale)
            else:
                self.cur_hysteresis -= 1
            self.last_overflow_iter = self.cur_iter
        else:
            if (self.cur_iter - self.last_overflow_iter) % self.scale_window == 0:
                self.cur_hysteresis = self.delayed_shift
                self.cur_scale *= 2.0
        self.cur_iter += 1


def sanitize_ds_config(cfg: dict) -> dict:
    """What the reference's ``_sanitize_ds_config`` enforces: an optimizer section must exist and be Adam(W)."""
    cfg = dict(cfg or {})
    opt = cfg.get("optimizer")
    if opt is None:
        raise ValueError("DeepSpeed config needs an 'optimizer' section (type Adam or AdamW).")
    if str(opt.get("type", "")).lower() not in ("adam", "adamw"):
        raise ValueError(f"Maggy only supports the Adam / AdamW optimizers with the DeepSpeed backend, got {opt.get('type')!r}.")
    return cfg


def get_maggy_deepspeed_engine(module: Type[torch.nn.Module], config_params: dict):
    """Class factory (same contract as the reference): instantiating the returned class builds ``module`` and returns an
    engine around it."""
    if getattr(module, "__name__", "") == "PipelineModule":
        raise AssertionError("Maggy does not support pipeline parallelism (reference modules.py:106-109)")
    cfg = sanitize_ds_config(config_params)
    from . import modules as pm

    stage = int(cfg.get("zero_optimization", {}).get("stage", 0))
    fp16 = dict(cfg.get("fp16", {}))
    mixed = bool(fp16.get("enabled") or cfg.get("bf16", {}).get("enabled"))
    inner_cls = pm.get_maggy_fsdp_wrapper(module, mixed) if stage >= 3 else pm.get_maggy_ddp_wrapper(module)
    user_module_path, user_module_name = module.__module__, module.__name__   # (`module` is a property name below)

    class MaggyDeepSpeedEngine(torch.nn.Module):
        __module__ = user_module_path

        def __init__(self, *args, **kwargs):
            super().__init__()
            self.wrapped = inner_cls(*args, **kwargs)
            if mixed and stage < 3 and not getattr(self.wrapped, "fused", False):
                self.wrapped = self.wrapped.to(torch.bfloat16)
            self.config = cfg
            self.zero_stage = stage
            self.world = dist.get_world_size() if dist.is_initialized() else 1
            self.gradient_accumulation_steps_ = int(cfg.get("gradient_accumulation_steps", 1))
            self.gradient_clipping_ = float(cfg.get("gradient_clipping", 0.0))
            micro = cfg.get("train_micro_batch_size_per_gpu")
            total = cfg.get("train_batch_size")
            if micro and total and int(total) != int(micro) * self.gradient_accumulation_steps_ * self.world:
                raise ValueError(f"train_batch_size {total} != micro batch {micro} x accumulation "
                                 f"{self.gradient_accumulation_steps_} x world {self.world}")
            static = float(fp16.get("loss_scale", 0) or 0)
            self.scaler: Optional[DynamicLossScaler] = None
            if fp16.get("enabled"):
                self.scaler = DynamicLossScaler(
                    init_scale=static if static > 0 else 2.0 ** int(fp16.get("initial_scale_power", 16)),
                    scale_window=fp16.get("loss_scale_window", 1000), hysteresis=fp16.get("hysteresis", 2),
                    min_scale=fp16.get("min_loss_scale", 1), dynamic=static <= 0)
            self.optimizer = self._make_optimizer(cfg["optimizer"])
            self.micro_steps = 0
            self.global_steps = 0
            self.skipped_steps = 0

        # ---------------------------------------------------------------- construction helpers
        @property
        def module(self):
            inner = getattr(self.wrapped, "inner", self.wrapped)
            return getattr(inner, "module", inner)

        def _flat(self):
            inner = getattr(self.wrapped, "inner", None)
            return getattr(inner, "flat", None)

        def _make_optimizer(self, section: dict):
            p = dict(section.get("params", {}))
            kw = {"lr": p.get("lr", 1e-3), "betas": tuple(p.get("betas", (0.9, 0.999))), "eps": p.get("eps", 1e-8),
                  "weight_decay": p.get("weight_decay", 0.0 if str(section["type"]).lower() == "adam" else 0.01)}
            flat = self._flat()
            if flat is not None:
                from ...parallel.arena import FusedAdamW

                return FusedAdamW(flat, max_grad_norm=self.gradient_clipping_, **kw)
            cls = torch.optim.AdamW if str(section["type"]).lower() == "adamw" else torch.optim.Adam
            return cls(self.wrapped.parameters(), **kw)

        # ---------------------------------------------------------------- engine API
        def forward(self, *args, **kwargs):
            if mixed:
                # 16-bit engines hold 16-bit parameters: floating-point inputs follow them (DeepSpeed users cast by
                # hand with .half(); doing it here keeps fp32 data loaders working)
                dt = next(self.wrapped.parameters()).dtype

                def cast(t):
                    return t.to(dt) if torch.is_tensor(t) and t.is_floating_point() and t.dtype != dt else t

                args = tuple(cast(a) for a in args)
                kwargs = {k: cast(v) for k, v in kwargs.items()}
            return self.wrapped(*args, **kwargs)

        @property
        def loss_scale(self) -> float:
            return self.scaler.cur_scale if self.scaler is not None else 1.0

        def gradient_accumulation_steps(self) -> int:
            return self.gradient_accumulation_steps_

        def is_gradient_accumulation_boundary(self) -> bool:
            return (self.micro_steps + 1) % self.gradient_accumulation_steps_ == 0

        def backward(self, loss, retain_graph: bool = False):
            scaled = loss.float() * (self.loss_scale / self.gradient_accumulation_steps_)
            scaled.backward(retain_graph=retain_graph)
            return loss

        def _overflow(self) -> bool:
            flat = self._flat()
            if flat is not None:
                flat.adopt_autograd_grads()
                bad = (~torch.isfinite(flat.grads)).any().to(torch.int32)
            else:
                grads = [p.grad for p in self.wrapped.parameters() if p.grad is not None]
                bad = torch.stack([(~torch.isfinite(g)).any() for g in grads]).any().to(torch.int32) if grads \
                    else torch.zeros((), dtype=torch.int32)
            if self.world > 1:
                dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            return bool(bad.item())

        def step(self):
            boundary = self.is_gradient_accumulation_boundary()
            self.micro_steps += 1
            if not boundary:
                return
            overflow = self._overflow() if self.scaler is not None else False
            flat = self._flat()
            if overflow:
                self.skipped_steps += 1
            else:
                if flat is not None:
                    self.optimizer.grad_scale = 1.0 / (self.world * self.loss_scale)
                    self.optimizer.step()
                else:
                    inv = 1.0 / self.loss_scale
                    params = [p for p in self.wrapped.parameters() if p.grad is not None]
                    if inv != 1.0:
                        for p in params:
                            p.grad.mul_(inv)
                    if self.gradient_clipping_ > 0:
                        torch.nn.utils.clip_grad_norm_(params, self.gradient_clipping_)
                    self.optimizer.step()
                self.global_steps += 1
            if self.scaler is not None:
                self.scaler.update(overflow)
            self.optimizer.zero_grad()

        def zero_grad(self):
            self.optimizer.zero_grad()

        def get_lr(self):
            return [g["lr"] for g in self.optimizer.param_groups]

    MaggyDeepSpeedEngine.__name__ = f"MaggyDeepSpeed_{user_module_name}"
    return MaggyDeepSpeedEngine
