"""Run a handful of launches of the tcgen05 GEMM on one training shape (target of `ncu --set full`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402

M, N, K = 32768, 3072, 768
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    F.gemm(a, b, out=out, block_n=256)
# wgrad shape (both operands MN-major)
dy = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
dw = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    F.gemm(dy, a, a_major=1, b_major=1, out=dw)
torch.cuda.synchronize()
