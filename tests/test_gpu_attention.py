"""K10 fused attention kernels vs a plain PyTorch fp32 reference (needs a B200)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, H, KVH, causal):
    B, T, _ = qkv.shape
    q, k, v = qkv.float().split([H * 64, KVH * 64, KVH * 64], dim=-1)
    q = q.view(B, T, H, 64).transpose(1, 2)
    k = k.view(B, T, KVH, 64).transpose(1, 2).repeat_interleave(H // KVH, dim=1)
    v = v.view(B, T, KVH, 64).transpose(1, 2).repeat_interleave(H // KVH, dim=1)
    s = q @ k.transpose(-1, -2) / 8.0
    if causal:
        s = s.masked_fill(torch.ones(T, T, device=qkv.device, dtype=torch.bool).triu(1), float("-inf"))
    lse2 = torch.logsumexp(s, dim=-1) / math.log(2.0)
    o = torch.softmax(s, dim=-1) @ v
    return o.transpose(1, 2).reshape(B, T, H * 64), lse2


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


CASES = [
    # B, T, H, KVH, causal
    (2, 128, 2, 2, True),
    (2, 256, 3, 3, True),
    (1, 1024, 4, 4, True),
    (2, 384, 4, 2, True),      # grouped-query
    (2, 256, 2, 2, False),
    (2, 200, 2, 2, True),      # ragged tail
    (1, 77, 2, 1, False),
]


@pytest.mark.parametrize("case", CASES)
def test_attention_fwd(case):
    from maggy_b200.ops import functional as F

    B, T, H, KVH, causal = case
    torch.manual_seed(0)
    qkv = torch.randn(B, T, (H + 2 * KVH) * 64, device="cuda").to(torch.bfloat16)
    o, lse = F.attention_fwd(qkv, H, KVH, causal=causal)
    o_ref, lse_ref = _ref(qkv, H, KVH, causal)
    assert _rel(o, o_ref) < 1e-2
    assert (lse - lse_ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("case", CASES)
def test_attention_bwd(case):
    from maggy_b200.ops import functional as F

    B, T, H, KVH, causal = case
    torch.manual_seed(1)
    qkv = torch.randn(B, T, (H + 2 * KVH) * 64, device="cuda").to(torch.bfloat16).requires_grad_(True)
    do = torch.randn(B, T, H * 64, device="cuda").to(torch.bfloat16)
    o = F.attention(qkv, H, KVH, causal=causal)
    o.backward(do)
    ref_in = qkv.detach().float().requires_grad_(True)
    o_ref, _ = _ref(ref_in, H, KVH, causal)
    o_ref.backward(do.float())
    g, g_ref = qkv.grad, ref_in.grad
    Cq, Ckv = H * 64, KVH * 64
    assert _rel(o, o_ref) < 1e-2
    assert _rel(g[..., :Cq], g_ref[..., :Cq]) < 2e-2, "dQ"
    assert _rel(g[..., Cq:Cq + Ckv], g_ref[..., Cq:Cq + Ckv]) < 2e-2, "dK"
    assert _rel(g[..., Cq + Ckv:], g_ref[..., Cq + Ckv:]) < 2e-2, "dV"


def test_gpt2_native_attention_matches_unfused_math():
    from maggy_b200.models.gpt2 import GPT2, GPT2Config

    torch.manual_seed(5)
    cfg = GPT2Config(vocab_size=1000, block_size=256, n_layer=2, n_head=2, n_embd=128)
    model = GPT2(cfg).cuda().to(torch.bfloat16)
    idx = torch.randint(0, 1000, (4, 256), device="cuda")
    losses, grads = [], []
    for impl in ("math", "native"):
        for blk in model.blocks:
            blk.attn_impl = impl
        model.zero_grad(set_to_none=True)
        loss = model(idx, idx)
        loss.backward()
        losses.append(loss.item())
        grads.append(model.blocks[0].qkv_w.grad.float().clone())
    assert abs(losses[0] - losses[1]) < 2e-2 * abs(losses[0])
    assert _rel(grads[1], grads[0]) < 5e-2
