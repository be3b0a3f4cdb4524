"""K10 attention vs cuDNN/flash SDPA (torch) on the GPT-2 small shape. CUDA events, L2 flushed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402
from bench_gemm import timeit  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for B, T, H in [(32, 1024, 12), (64, 1024, 12), (8, 4096, 16)]:
    qkv = torch.randn(B, T, 3 * H * 64, device=dev).to(torch.bfloat16)
    q, k, v = (t.transpose(1, 2) for t in qkv.view(B, T, 3, H, 64).unbind(2))
    t_ours = timeit(lambda: F.attention_fwd(qkv, H, causal=True), flush)
    t_ref = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True), flush)
    fl = 4.0 * B * H * T * T * 64 / 2
    line = (f"fwd B={B} T={T} H={H}: ours {t_ours:7.3f} ms ({fl / t_ours / 1e9:5.0f} TF)  "
            f"sdpa {t_ref:7.3f} ms ({fl / t_ref / 1e9:5.0f} TF)  speedup {t_ref / t_ours:.2f}")
    if hasattr(F, "attention"):
        qkv_g = qkv.clone().requires_grad_(True)
        qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        do = torch.randn(B, T, H * 64, device=dev).to(torch.bfloat16)
        do_t = do.view(B, T, H, 64).transpose(1, 2)

        def ours_fb():
            qkv_g.grad = None
            F.attention(qkv_g, H, causal=True).backward(do)

        def ref_fb():
            qg.grad = kg.grad = vg.grad = None
            torch.nn.functional.scaled_dot_product_attention(qg, kg, vg, is_causal=True).backward(do_t)

        t_of, t_rf = timeit(ours_fb, flush), timeit(ref_fb, flush)
        line += f" | fwd+bwd ours {t_of:7.3f} ms sdpa {t_rf:7.3f} ms speedup {t_rf / t_of:.2f}"
    print(line, flush=True)
