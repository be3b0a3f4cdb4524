"""Per-node cost of a CUDA graph made of dependent tiny kernels on this GPU: N launches of a 1-CTA kernel captured in
one graph, replay time / N. MAGGY_B200_PDL=0/1 toggles the programmatic-dependent-launch attribute of our launcher."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.ops import batchnorm as B  # noqa: E402
from maggy_b200.ops import functional as F  # noqa: E402

lib = B._cuda()
dev = "cuda:0"
C = 64
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
coef = torch.empty(2, C, device=dev)
N = 400


def chain():
    s = F._stream()
    for _ in range(N):
        lib.mg_bn_eval_coef(rm.data_ptr(), rv.data_ptr(), None, None, 0, 1e-5, C, coef.data_ptr(), s)


def aten_chain():
    for _ in range(N):
        coef.add_(1.0)


for name, fn in (("maggy_b200 1-CTA kernel", chain), ("ATen elementwise add_", aten_chain)):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s}: {e0.elapsed_time(e1) / 10 / N * 1e3:6.2f} us per graph node (PDL={os.environ.get('MAGGY_B200_PDL', '1')})")
