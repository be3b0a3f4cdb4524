"""ResNet-50 training step (batch 64, bf16 channels_last, synthetic data) as the ASHA trial body runs it:
`build_trainer` = flat arena + fused AdamW + the whole step as one CUDA graph. Arms: RESNET_ARM=graph (native
convolutions; MAGGY_B200_BN=native|aten selects the norm kernels) or eager (cuDNN + ATen + torch SGD, op by op)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.models.resnet import resnet50  # noqa: E402
from maggy_b200.parallel.trainer import build_trainer  # noqa: E402

B = int(os.environ.get("RESNET_B", 64))
arm = os.environ.get("RESNET_ARM", "graph")
iters = int(os.environ.get("RESNET_ITERS", 30))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(B, 3, 224, 224, device=dev).to(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (B,), device=dev)
if arm == "graph":
    with torch.device(dev):
        model = resnet50(1000)
    model = model.to(memory_format=torch.channels_last)

    def loss_fn(m, xx, yy):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return torch.nn.functional.cross_entropy(m(xx), yy)

    _, _, step = build_trainer(model, (B, 3, 224, 224), dev, target_shape=(B,), memory_format=torch.channels_last,
                               loss_fn=loss_fn, lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-4, warmup_iters=2)
    step.load_batch(x, y)
    fn = step
else:
    model = resnet50(1000, conv_impl="torch").to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True)

    def fn():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
        return loss

for _ in range(5):
    loss = fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    loss = fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(json.dumps({"arm": arm, "bn": os.environ.get("MAGGY_B200_BN", "native"), "batch": B, "ms_per_step": ms,
                  "img_per_s": B / ms * 1e3, "loss": float(loss)}))
