"""Bandwidth-bound kernels vs the measured copy roofline (MEASURED_PEAKS.json: 6576 GB/s). CUDA events, L2 flushed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402
from bench_gemm import timeit  # noqa: E402

dev = "cuda:0"
PEAK = 6576e9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_copy_gbps"] * 1e9
except Exception:  # noqa: BLE001
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
M, C, V = 65536, 768, 50304
rows = []


def report(name, t_ms, bytes_moved):
    roof = bytes_moved / PEAK * 1e3
    rows.append({"kernel": name, "ms": t_ms, "roofline_ms": roof, "fraction_of_roofline": roof / t_ms})
    print(f"{name:44s} {t_ms:8.3f} ms   roofline {roof:7.3f} ms   {roof / t_ms * 100:5.1f} %", flush=True)


x = torch.randn(M, C, device=dev).to(torch.bfloat16)
w = torch.ones(C, device=dev, dtype=torch.bfloat16)
b = torch.zeros(C, device=dev, dtype=torch.bfloat16)
report("layer_norm fwd [65536 x 768]", timeit(lambda: F.layer_norm_fwd(x, w, b), flush), 2 * M * C * 2)
y, mean, rstd = F.layer_norm_fwd(x, w, b)
dy = torch.randn_like(x)
dg, db = torch.zeros(C, device=dev, dtype=torch.bfloat16), torch.zeros(C, device=dev, dtype=torch.bfloat16)
report("layer_norm bwd [65536 x 768]", timeit(lambda: F.layer_norm_bwd(dy, x, w, mean, rstd, dg, db), flush), 3 * M * C * 2)
rg = torch.randn_like(x)
dcol = torch.zeros(C, device=dev, dtype=torch.bfloat16)
report("layer_norm bwd + residual grad [65536 x 768]",
       timeit(lambda: F.layer_norm_bwd(dy, x, w, mean, rstd, dg, db, res_grad=rg), flush), 4 * M * C * 2)
report("layer_norm bwd + residual + dx column sums",
       timeit(lambda: F.layer_norm_bwd(dy, x, w, mean, rstd, dg, db, res_grad=rg, dcol=dcol), flush), 4 * M * C * 2)
if os.environ.get("BW_ONLY_NORM"):
    sys.exit(0)
for n in (768, 3072):
    d = torch.randn(M, n, device=dev).to(torch.bfloat16)
    out = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    report(f"colsum (bias grad) [65536 x {n}]", timeit(lambda: F.colsum(d, out), flush), M * n * 2)
h = torch.randn(M, 4 * C, device=dev).to(torch.bfloat16)
report("gelu fwd [65536 x 3072]", timeit(lambda: F.gelu_fwd(h), flush), 2 * M * 4 * C * 2)
Mc = 16384  # cross entropy on a quarter of the rows (3.3 GB logits would not leave room for the flush buffer twice)
logits = torch.randn(Mc, V, device=dev).to(torch.bfloat16)
tgt = torch.randint(0, 50257, (Mc,), device=dev)
work = logits.clone()


def ce():
    work.copy_(logits)
    return F.cross_entropy(work, tgt, v_real=50257)


t_copy = timeit(lambda: work.copy_(logits), flush)
t_ce = timeit(ce, flush) - t_copy
report(f"cross_entropy fwd+bwd [{Mc} x 50304] (copy subtracted)", t_ce, 2 * Mc * V * 2)
report("  (device copy of the same logits)", t_copy, 2 * Mc * V * 2)
n_par = 124_475_904
master = torch.zeros(n_par, device=dev)
m1, v1 = torch.zeros_like(master), torch.zeros_like(master)
g = torch.zeros(n_par, device=dev, dtype=torch.bfloat16)
pb = torch.zeros(n_par, device=dev, dtype=torch.bfloat16)
report("fused AdamW, 124.5 M params", timeit(lambda: F.adam_flat_step(
    master, m1, v1, g, pb, decay_boundary=n_par, lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, step=1,
    grad_scale=1.0), flush), n_par * 28)
qkv = torch.randn(M, 3 * C, device=dev).to(torch.bfloat16)
q8, s8 = F.quantize_fp8(x)
report("quantize e4m3 [65536 x 768]", timeit(lambda: F.quantize_fp8(x), flush), M * C * (2 + 2 + 1))
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
