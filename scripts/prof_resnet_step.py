"""One eager ResNet-50 training step (batch 64, channels_last bf16 autocast) for `ncu --metrics gpu__time_duration.sum`:
launch list of the ASHA trial body. PROF_IMPL=native|torch selects the convolution implementation."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.models.resnet import resnet50  # noqa: E402

impl = os.environ.get("PROF_IMPL", "native")
B = int(os.environ.get("PROF_B", 64))
dev = "cuda:0"
torch.manual_seed(0)
model = resnet50(1000, conv_impl=impl).to(dev).to(memory_format=torch.channels_last)
opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True)
x = torch.randn(B, 3, 224, 224, device=dev).to(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (B,), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(model(x), y)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
