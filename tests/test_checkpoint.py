"""Flat-arena checkpoint round trip (CPU: save / load do not need the kernels)."""
import torch

from maggy_b200.parallel.arena import FlatParams, FusedAdamW
from maggy_b200.utils import checkpoint


def _make():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 8))
    flat = FlatParams(m, device="cpu")
    return m, flat, FusedAdamW(flat, lr=1e-3)


def test_checkpoint_round_trip(tmp_path):
    m, flat, opt = _make()
    opt.step_count = 17
    opt.m.uniform_(-1, 1)
    opt.v.uniform_(0, 1)
    flat.master.uniform_(-1, 1)
    flat.params.copy_(flat.master.to(torch.bfloat16))
    d = checkpoint.save_training_state(str(tmp_path / "ck" / "17"), flat, opt, extra={"epoch": 3})
    m2, flat2, opt2 = _make()
    extra = checkpoint.load_training_state(d, flat2, opt2)
    assert extra == {"epoch": 3} and opt2.step_count == 17
    assert torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v) and torch.equal(flat2.master, flat.master)
    assert torch.equal(flat2.params, flat.params)
    assert torch.equal(m2[0].weight, m[0].weight)      # module parameters are views of the arena
    assert checkpoint.latest(str(tmp_path / "ck")) == d


def test_checkpoint_rejects_a_different_model(tmp_path):
    import pytest

    _, flat, opt = _make()
    d = checkpoint.save_training_state(str(tmp_path / "c"), flat, opt)
    other = torch.nn.Sequential(torch.nn.Linear(16, 64))
    flat2 = FlatParams(other, device="cpu")
    with pytest.raises(ValueError):
        checkpoint.load_training_state(d, flat2, FusedAdamW(flat2, lr=1e-3))
