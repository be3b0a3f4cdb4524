"""fp8 (e4m3) tcgen05 GEMM + device-side quantisation vs plain PyTorch fp32 references (needs a B200)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("shape", [(1000, 96), (4096, 1024), (7, 8), (3, 11)])
def test_quantize_fp8(shape):
    from maggy_b200.ops import functional as F

    torch.manual_seed(0)
    x = (torch.randn(*shape, device="cuda") * 3).to(torch.bfloat16)
    q, scale = F.quantize_fp8(x)
    amax = x.float().abs().max()
    assert abs(scale.item() - amax.item() / 448.0) < 1e-6 * max(1.0, amax.item())
    ref = (x.float() / scale).clamp(-448, 448).to(torch.float8_e4m3fn)
    # same rounding as the ATen conversion up to ties: compare the dequantised values
    deq, deq_ref = q.float() * scale, ref.float() * scale
    assert _rel(deq, deq_ref) < 2e-2
    assert _rel(deq, x) < 5e-2


@pytest.mark.parametrize("mnk", [(256, 256, 256), (512, 768, 1024), (4096, 3072, 768), (300, 264, 128), (128, 8, 16),
                                 (1000, 1000, 4096)])
@pytest.mark.parametrize("bias", [False, True])
def test_gemm_fp8(mnk, bias):
    from maggy_b200.ops import functional as F

    M, N, K = mnk
    torch.manual_seed(1)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bv = torch.randn(N, device="cuda").to(torch.bfloat16) if bias else None
    aq, sa = F.quantize_fp8(a)
    bq, sb = F.quantize_fp8(b)
    y = F.gemm_fp8(aq, bq, sa, sb, bias=bv)
    ref = (aq.float() * sa) @ (bq.float() * sb).t()
    if bias:
        ref = ref + bv.float()
    assert _rel(y, ref) < 6e-3  # only the bf16 rounding of the output separates them
    # and the quantised product tracks the bf16 one
    assert _rel(y, a.float() @ b.float().t() + (bv.float() if bias else 0)) < 6e-2


def test_linear_fp8_autograd():
    from maggy_b200.ops import functional as F

    torch.manual_seed(2)
    x = torch.randn(4, 64, 256, device="cuda").to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(512, 256, device="cuda") / 16).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(512, device="cuda").to(torch.bfloat16).requires_grad_(True)
    y = F.linear_fp8(x, w, b)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.linear(xr, wr, br)
    assert _rel(y, yr) < 6e-2
    g = torch.randn_like(yr)
    y.backward(g.to(torch.bfloat16))
    yr.backward(g.to(torch.bfloat16).float())
    assert _rel(x.grad, xr.grad) < 2e-2 and _rel(w.grad, wr.grad) < 2e-2 and _rel(b.grad, br.grad) < 2e-2


def test_llama_fp8_forward_backward():
    from maggy_b200.models.llama import Llama, LlamaConfig

    torch.manual_seed(3)
    cfg = LlamaConfig.named("llama-tiny")
    m = Llama(cfg).cuda().to(torch.bfloat16)
    idx = torch.randint(0, cfg.vocab_size, (4, 64), device="cuda")
    loss_bf16 = m(idx, idx)
    cfg.fp8 = True
    loss_fp8 = m(idx, idx)
    loss_fp8.backward()
    assert torch.isfinite(loss_fp8)
    assert abs(loss_fp8.item() - loss_bf16.item()) < 0.05 * abs(loss_bf16.item())
    assert all(torch.isfinite(p.grad.float()).all() for p in m.parameters() if p.grad is not None)
