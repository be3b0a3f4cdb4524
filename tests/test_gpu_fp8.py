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


# ---------------------------------------------------------------------------------------------------------------
# block-scaled fp8 (MXFP8: e4m3 + one UE8M0 scale per 32 K-elements, tcgen05.mma.kind::mxf8f6f4.block_scale)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(128, 128), (300, 1024), (1000, 96), (5, 48)])
def test_quantize_mxfp8(shape):
    from maggy_b200.ops import functional as F

    torch.manual_seed(3)
    rows, K = shape
    # rows with very different magnitudes: a per-tensor scale would flush the small ones, block scales keep them
    x = (torch.randn(rows, K, device="cuda") * torch.logspace(-3, 2, rows, device="cuda").view(-1, 1)).to(torch.bfloat16)
    q, sf = F.quantize_mxfp8(x)
    deq = F.dequantize_mxfp8(q, sf)
    assert q.shape == x.shape and sf.numel() == ((rows + 127) // 128) * ((K + 127) // 128) * 512
    assert _rel(deq, x) < 4e-2
    # every 32-block's scale is the smallest power of two that brings its amax under 448
    blocks = x.float().view(rows, -1, 32) if K % 32 == 0 else None
    if blocks is not None:
        amax = blocks.abs().amax(-1)
        scale = (deq.view(rows, -1, 32).abs().amax(-1) / q.float().view(rows, -1, 32).abs().amax(-1).clamp_min(1e-30))
        ok = amax > 0
        assert torch.all((amax[ok] / scale[ok]) <= 448.0 * 1.001)
        assert torch.all((amax[ok] / scale[ok]) > 224.0 * 0.999)
    # small rows survive (per-row relative error, not just the global norm)
    row_err = (deq - x.float()).norm(dim=1) / x.float().norm(dim=1).clamp_min(1e-30)
    assert row_err.max().item() < 6e-2


@pytest.mark.parametrize("mnk", [(128, 128, 128), (256, 384, 512), (4096, 3072, 768), (300, 264, 160), (1000, 1000, 4096)])
@pytest.mark.parametrize("bias", [False, True])
def test_gemm_mxfp8(mnk, bias):
    from maggy_b200.ops import functional as F

    M, N, K = mnk
    torch.manual_seed(4)
    a = (torch.randn(M, K, device="cuda") * torch.logspace(-2, 1, M, device="cuda").view(-1, 1)).to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bv = torch.randn(N, device="cuda").to(torch.bfloat16) if bias else None
    aq, sfa = F.quantize_mxfp8(a)
    bq, sfb = F.quantize_mxfp8(b)
    y = F.gemm_mxfp8(aq, sfa, bq, sfb, bias=bv)
    ref = F.dequantize_mxfp8(aq, sfa) @ F.dequantize_mxfp8(bq, sfb).t()
    if bias:
        ref = ref + bv.float()
    assert _rel(y, ref) < 6e-3      # exact block-scaled product up to fp32 accumulation order + bf16 output rounding
    assert _rel(y, a.float() @ b.float().t() + (bv.float() if bias else 0)) < 6e-2


def test_linear_mxfp8_autograd():
    from maggy_b200.ops import functional as F

    torch.manual_seed(5)
    x = torch.randn(4, 64, 256, device="cuda").to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(512, 256, device="cuda") / 16).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(512, device="cuda").to(torch.bfloat16).requires_grad_(True)
    y = F.linear_mxfp8(x, w, b)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.linear(xr, wr, br)
    assert _rel(y, yr) < 6e-2
    g = torch.randn_like(y)
    y.backward(g)
    yr.backward(g.float())
    assert _rel(x.grad, xr.grad) < 2e-2 and _rel(w.grad, wr.grad) < 2e-2 and _rel(b.grad, br.grad) < 2e-2
