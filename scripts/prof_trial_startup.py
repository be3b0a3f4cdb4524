"""Per-trial start-up cost of the ASHA ResNet-50 trial body inside one worker process: consecutive trials each build
a model, a flat arena, a fused optimizer and capture a fresh step graph, then run 100 steps."""
import gc
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.models.resnet import resnet50  # noqa: E402
from maggy_b200.parallel.trainer import build_trainer  # noqa: E402

dev = torch.device("cuda", 0)
B = 64
x = torch.randn(B, 3, 224, 224, device=dev).to(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (B,), device=dev)


def loss_fn(m, xx, yy):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return torch.nn.functional.cross_entropy(m(xx), yy)


def sync():
    torch.cuda.synchronize()
    return time.time()


for trial in range(6):
    t0 = sync()
    with torch.device(dev):
        model = resnet50(1000, dropout=0.1 * (trial % 3))
    model = model.to(memory_format=torch.channels_last)
    t1 = sync()
    _, _, step = build_trainer(model, (B, 3, 224, 224), dev, target_shape=(B,), memory_format=torch.channels_last,
                               loss_fn=loss_fn, lr=1e-3, warmup_iters=2)
    step.load_batch(x, y)
    t2 = sync()
    step()
    t3 = sync()
    for _ in range(99):
        step()
    t4 = sync()
    del model, step
    gc.collect()
    t5 = sync()
    print(f"trial {trial}: model {t1 - t0:.3f}s  arena+optimizer {t2 - t1:.3f}s  warm-up+capture {t3 - t2:.3f}s  "
          f"99 steps {t4 - t3:.3f}s  teardown {t5 - t4:.3f}s  | pool={os.environ.get('MAGGY_B200_GRAPH_POOL', 'shared')}",
          flush=True)
