"""FlatParams adoption of autograd-produced gradients (CPU logic; the fused optimizer itself needs a GPU)."""
import torch

from maggy_b200.parallel.arena import FlatParams


def test_adopt_autograd_grads_accumulate_in_place():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.BatchNorm1d(8), torch.nn.Linear(8, 4)).to(torch.bfloat16)
    flat = FlatParams(m, device="cpu")
    x = torch.randn(4, 8).to(torch.bfloat16)
    flat.start_step()
    m(x).float().sum().backward()
    flat.adopt_autograd_grads()
    assert flat._autograd_ranges == [(0, flat.numel)]
    for _, p, _, _ in flat.entries:
        assert p.grad.data_ptr() == p.main_grad.data_ptr()
    g1 = flat.grads.clone()
    assert g1.float().abs().sum() > 0
    flat.start_step()
    assert flat.grads.float().abs().sum() == 0
    flat.start_step()  # idempotent
    m(x).float().sum().backward()  # AccumulateGrad adds into the arena slice
    flat.adopt_autograd_grads()
    assert torch.equal(flat.grads, g1)
