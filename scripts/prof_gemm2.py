"""ncu target: our GEMM (whatever MAGGY_B200_GEMM_* selects) and cuBLAS on the fwd fc shape, for side-by-side metrics."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402

M, N, K = 32768, 3072, 768
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    F.gemm(a, b, out=out, block_n=256)
for _ in range(4):
    torch.matmul(a, b.t(), out=out)
torch.cuda.synchronize()
