"""K10 attention vs cuDNN/flash SDPA (torch). CUDA events, L2 flushed, median of 10.

  python scripts/bench_attn.py [--shapes 64x1024x12,8x4096x16] [--no-ref] [--tag NAME] [--json out.json]

Kernel knobs are environment variables read once per process (csrc/ops/attention_v2.cu): MAGGY_B200_ATTN_GEN,
MAGGY_B200_ATTN_POLY, MAGGY_B200_ATTN_POLY_BWD, MAGGY_B200_ATTN_BWD_ORDER."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from maggy_b200.ops import functional as F  # noqa: E402
from bench_gemm import timeit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="32x1024x12,64x1024x12,8x4096x16")
ap.add_argument("--no-ref", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--json", default="")
args = ap.parse_args()

dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
knobs = {k: v for k, v in os.environ.items() if k.startswith("MAGGY_B200_ATTN")}
rows = []
for shp in args.shapes.split(","):
    B, T, H = (int(x) for x in shp.split("x"))
    qkv = torch.randn(B, T, 3 * H * 64, device=dev).to(torch.bfloat16)
    q, k, v = (t.transpose(1, 2) for t in qkv.view(B, T, 3, H, 64).unbind(2))
    fl = 4.0 * B * H * T * T * 64 / 2
    rec = {"B": B, "T": T, "H": H, "tag": args.tag, "knobs": knobs}
    rec["fwd_ms"] = timeit(lambda: F.attention_fwd(qkv, H, causal=True), flush)
    qkv_g = qkv.clone().requires_grad_(True)
    do = torch.randn(B, T, H * 64, device=dev).to(torch.bfloat16)

    def ours_fb():
        qkv_g.grad = None
        F.attention(qkv_g, H, causal=True).backward(do)

    rec["fwdbwd_ms"] = timeit(ours_fb, flush)
    line = (f"[{args.tag}] B={B} T={T} H={H}: fwd {rec['fwd_ms']:7.3f} ms ({fl / rec['fwd_ms'] / 1e9:5.0f} TF)  "
            f"fwd+bwd {rec['fwdbwd_ms']:7.3f} ms ({3.5 * fl / rec['fwdbwd_ms'] / 1e9:5.0f} TF)")
    if not args.no_ref:
        qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        do_t = do.view(B, T, H, 64).transpose(1, 2)

        def ref_fb():
            qg.grad = kg.grad = vg.grad = None
            torch.nn.functional.scaled_dot_product_attention(qg, kg, vg, is_causal=True).backward(do_t)

        rec["sdpa_fwd_ms"] = timeit(
            lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True), flush)
        rec["sdpa_fwdbwd_ms"] = timeit(ref_fb, flush)
        line += (f" | sdpa fwd {rec['sdpa_fwd_ms']:7.3f} fwd+bwd {rec['sdpa_fwdbwd_ms']:7.3f} | speedup fwd "
                 f"{rec['sdpa_fwd_ms'] / rec['fwd_ms']:.2f} fwd+bwd {rec['sdpa_fwdbwd_ms'] / rec['fwdbwd_ms']:.2f}")
    print(line, flush=True)
    rows.append(rec)
if args.json:
    with open(args.json, "a") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
