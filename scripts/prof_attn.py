import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402

B, T, H = 64, 1024, 12
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").to(torch.bfloat16)
for _ in range(4):
    o, lse = F.attention_fwd(qkv, H, causal=True)
if hasattr(F, "attention"):
    qkv.requires_grad_(True)
    do = torch.randn(B, T, H * 64, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        qkv.grad = None
        F.attention(qkv, H, causal=True).backward(do)
torch.cuda.synchronize()
