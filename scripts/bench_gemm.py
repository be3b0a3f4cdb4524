"""Micro-benchmark: tcgen05 GEMM (ours) vs torch.matmul (cuBLAS) on the GPT-2 small training shapes.
CUDA-event timed, L2 flushed between iterations. Usage: python scripts/bench_gemm.py [--tokens 32768]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402


def timeit(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=32768)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = "cuda:0"
    M = args.tokens
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    C = 768
    shapes = [  # (name, M, N, K, a_major, b_major)
        ("fwd qkv      x[M,768] W[2304,768]", M, 3 * C, C, 0, 0),
        ("fwd proj     x[M,768] W[768,768]", M, C, C, 0, 0),
        ("fwd fc       x[M,768] W[3072,768]", M, 4 * C, C, 0, 0),
        ("fwd fc2      x[M,3072] W[768,3072]", M, C, 4 * C, 0, 0),
        ("fwd lm_head  x[M,768] W[50304,768]", M, 50304, C, 0, 0),
        ("dgrad fc     dy[M,3072] W[3072,768]", M, C, 4 * C, 0, 1),
        ("dgrad lm     dy[M,50304] W[50304,768]", M, C, 50304, 0, 1),
        ("wgrad fc     dy[M,3072]^T x[M,768]", 4 * C, C, M, 1, 1),
        ("wgrad qkv    dy[M,2304]^T x[M,768]", 3 * C, C, M, 1, 1),
        ("wgrad lm     dy[M,50304]^T x[M,768]", 50304, C, M, 1, 1),
        ("square 8192", 8192, 8192, 8192, 0, 0),
    ]
    rows = []
    for name, m, n, k, am, bm in shapes:
        a = torch.randn((m, k) if am == 0 else (k, m), device=dev, dtype=torch.bfloat16)
        b = torch.randn((n, k) if bm == 0 else (k, n), device=dev, dtype=torch.bfloat16)
        out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        ta = a if am == 0 else a.t()
        tb = b.t() if bm == 0 else b
        res = {}
        for bn in (0, 192, 256):
            try:
                res[bn] = timeit(lambda: F.gemm(a, b, a_major=am, b_major=bm, out=out, block_n=bn), flush)
            except Exception as e:  # noqa: BLE001
                res[bn] = float("nan")
                print("ERR", name, bn, e)
        t_ref = timeit(lambda: torch.matmul(ta, tb, out=out), flush)
        fl = 2.0 * m * n * k
        row = {"shape": name, "M": m, "N": n, "K": k,
               "ours_auto_ms": res[0], "ours_bn192_ms": res[192], "ours_bn256_ms": res[256], "cublas_ms": t_ref,
               "ours_tflops": fl / (res[0] * 1e-3) / 1e12, "cublas_tflops": fl / (t_ref * 1e-3) / 1e12}
        rows.append(row)
        print(f"{name:42s} ours auto {res[0]:8.3f} ms ({row['ours_tflops']:7.1f} TF) bn192 {res[192]:8.3f} bn256 {res[256]:8.3f}"
              f"   cuBLAS {t_ref:8.3f} ms ({row['cublas_tflops']:7.1f} TF)  ratio {t_ref / res[0]:.2f}", flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
