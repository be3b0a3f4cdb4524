import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402

a = torch.randn(16384, 2048, device="cuda").to(torch.bfloat16)
b = torch.randn(16384, 2048, device="cuda").to(torch.bfloat16)
am, sfa = F.quantize_mxfp8(a)
bm, sfb = F.quantize_mxfp8(b)
for _ in range(4):
    y = F.gemm_mxfp8(am, sfa, bm, sfb)
torch.cuda.synchronize()
