"""ASHA hyperparameter search over ResNet-50 where every trial trains through ONE captured CUDA graph:
K7 implicit-GEMM convolutions + ATen BatchNorm adopted into the flat arena + the fused AdamW kernel.

    python examples/resnet_asha_graph.py            # needs >= 1 B200; one worker per visible GPU
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from maggy_b200 import Searchspace, experiment
from maggy_b200.config import HyperparameterOptConfig
from maggy_b200.optimizer import Asha


def train_fn(hparams, reporter):
    import torch

    from maggy_b200.models.resnet import resnet50
    from maggy_b200.parallel.trainer import build_trainer

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):                       # initialise on the GPU: trial set-up in milliseconds
        model = resnet50(1000, dropout=hparams["dropout"])
    model = model.to(memory_format=torch.channels_last)

    def loss_fn(m, x, y):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return torch.nn.functional.cross_entropy(m(x), y)

    B = 64
    _, _, step = build_trainer(model, (B, 3, 224, 224), dev, target_shape=(B,), memory_format=torch.channels_last,
                               loss_fn=loss_fn, lr=hparams["lr"], weight_decay=hparams["wd"], warmup_iters=2)
    x = torch.randn(B, 3, 224, 224, device=dev).to(memory_format=torch.channels_last)   # synthetic batch
    y = torch.randint(0, 1000, (B,), device=dev)
    step.load_batch(x, y)
    loss = None
    for i in range(int(hparams.get("budget", 1)) * 20):
        loss = step()                              # graph replay
        if i % 5 == 4:
            reporter.broadcast(-loss.detach().float(), i)   # device tensor: published by a kernel, no sync
    return -float(loss)


if __name__ == "__main__":
    sp = Searchspace(lr=("DOUBLE", [1e-4, 3e-3]), wd=("DOUBLE", [1e-5, 1e-2]), dropout=("DOUBLE", [0.0, 0.5]))
    cfg = HyperparameterOptConfig(num_trials=16, optimizer=Asha(reduction_factor=2, resource_min=1, resource_max=4),
                                  searchspace=sp, direction="max", es_policy="median", es_min=4, name="resnet_asha")
    print(experiment.lagom(train_fn, cfg))
