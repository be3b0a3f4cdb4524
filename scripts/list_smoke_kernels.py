"""Kernel names launched by __graft_entry__.smoke() (torch profiler, CUDA activities): none of them may be a library
attention kernel."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.smoke()
    torch.cuda.synchronize()
names = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        names[e.name] = names.get(e.name, 0) + 1
for n, c in sorted(names.items(), key=lambda kv: -kv[1]):
    print(f"{c:4d}  {n[:110]}")
bad = [n for n in names if "sdpa" in n.lower() or "flash" in n.lower() or "fmha" in n.lower()]
print("LIBRARY_ATTENTION_KERNELS:", bad)
