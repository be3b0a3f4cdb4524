"""A few launches of each kernel that got no ncu capture so far (target of `ncu --set full -k regex:...`):
MN-major dgrad / wgrad GEMMs at the B = 64 training shape, implicit-GEMM conv fprop / wgrad, the fused BatchNorm
kernels and the staged norm backward. PROF_WHAT selects the group so one ncu invocation stays short."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import functional as F  # noqa: E402

what = os.environ.get("PROF_WHAT", "gemm")
dev = "cuda:0"
if what == "gemm":
    M = 65536
    dy = torch.randn(M, 768, device=dev, dtype=torch.bfloat16)
    w2 = torch.randn(768, 3072, device=dev, dtype=torch.bfloat16)     # fc2 weight [out=768, in=3072]
    for _ in range(4):
        F.gemm(dy, w2, b_major=F.MN_MAJOR)                            # dgrad: [M,768] x [768,3072]
    dyf = torch.randn(M, 3072, device=dev, dtype=torch.bfloat16)
    x = torch.randn(M, 768, device=dev, dtype=torch.bfloat16)
    dw = torch.empty(3072, 768, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        F.gemm(dyf, x, a_major=F.MN_MAJOR, b_major=F.MN_MAJOR, out=dw)  # wgrad fc: dy^T x
elif what == "conv":
    from maggy_b200.ops import conv as C

    x = torch.randn(64, 14, 14, 256, device=dev, dtype=torch.bfloat16)
    w = torch.randn(256, 3, 3, 256, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(64, 14, 14, 256, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        C.conv2d_fprop(x, w, 1)
        C.conv2d_wgrad(x, dy, 3, 3, 1)
    x = torch.randn(64, 56, 56, 64, device=dev, dtype=torch.bfloat16)
    w = torch.randn(64, 3, 3, 64, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(64, 56, 56, 64, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        C.conv2d_fprop(x, w, 1)
        C.conv2d_wgrad(x, dy, 3, 3, 1)
else:
    from maggy_b200.ops.batchnorm import BatchNorm2d

    bn = BatchNorm2d(256, relu=True).to(dev)
    x = torch.randn(64, 256, 56, 56, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = torch.randn_like(x).requires_grad_(True)
    g = torch.randn_like(x)
    for _ in range(3):
        bn(x, r).backward(g)
    M, C_ = 65536, 768
    xx = torch.randn(M, C_, device=dev).to(torch.bfloat16)
    wn = torch.ones(C_, device=dev, dtype=torch.bfloat16)
    bnn = torch.zeros(C_, device=dev, dtype=torch.bfloat16)
    y, mean, rstd = F.layer_norm_fwd(xx, wn, bnn)
    dyy, rg = torch.randn_like(xx), torch.randn_like(xx)
    dg, db, dc = torch.zeros_like(wn), torch.zeros_like(wn), torch.zeros_like(wn)
    for _ in range(3):
        F.layer_norm_bwd(dyy, xx, wn, mean, rstd, dg, db, res_grad=rg, dcol=dc)
torch.cuda.synchronize()
