"""fp8 (e4m3, per-tensor) and block-scaled MXFP8 tcgen05 GEMMs vs our bf16 GEMM vs cuBLASLt fp8 (torch._scaled_mm).
CUDA events, L2 flushed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402
from bench_gemm import timeit  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name, M, N, K in [("fc 32768x3072x768", 32768, 3072, 768), ("fc2 32768x768x3072", 32768, 768, 3072),
                      ("llama w13 16384x16384x2048", 16384, 16384, 2048), ("square 8192", 8192, 8192, 8192)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    aq, sa = F.quantize_fp8(a)
    bq, sb = F.quantize_fp8(b)
    t8 = timeit(lambda: F.gemm_fp8(aq, bq, sa, sb, out=out), flush)
    t16 = timeit(lambda: F.gemm(a, b, out=out), flush)
    tq = timeit(lambda: F.quantize_fp8(a), flush)
    am, sfa = F.quantize_mxfp8(a)
    bm, sfb = F.quantize_mxfp8(b)
    tmx = timeit(lambda: F.gemm_mxfp8(am, sfa, bm, sfb, out=out), flush)
    tqm = timeit(lambda: F.quantize_mxfp8(a), flush)
    try:
        tl = timeit(lambda: torch._scaled_mm(aq, bq.t(), scale_a=sa, scale_b=sb, out_dtype=torch.bfloat16), flush)
    except Exception as e:  # noqa: BLE001
        tl = float("nan")
        print("scaled_mm:", str(e)[:100])
    fl = 2.0 * M * N * K
    print(f"{name:30s} fp8 ours {t8:7.3f} ms ({fl / t8 / 1e9:6.0f} TF)  bf16 ours {t16:7.3f} ms ({fl / t16 / 1e9:6.0f} TF)"
          f"  cuBLASLt fp8 {tl:7.3f} ms ({fl / tl / 1e9:6.0f} TF)  quantise(A) {tq:6.3f} ms"
          f" | mxfp8 (block-scaled) ours {tmx:7.3f} ms ({fl / tmx / 1e9:6.0f} TF)  mx quantise(A) {tqm:6.3f} ms", flush=True)
