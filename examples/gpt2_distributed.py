#!/usr/bin/env python
"""Distributed GPT-2 training through ``experiment.lagom`` + ``TorchDistributedConfig``: the train_fn is
oblivious of the number of GPUs; the framework starts one worker per GPU, builds the process group, and the
``module`` class it hands in makes the model data parallel on the fused NVLink path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def train_fn(module, hparams, reporter):
    import torch

    from maggy_b200.models.gpt2 import GPT2Config

    cfg = GPT2Config.named(hparams["model"])
    model = module(cfg)                                   # built, sharded and wrapped by the framework
    opt = torch.optim.AdamW(model.parameters(), lr=hparams["lr"], weight_decay=0.1)   # -> FusedAdamW under zero_lvl>=1
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(int(os.environ.get("RANK", "0")))
    loss = None
    for step in range(hparams["steps"]):
        idx = torch.randint(0, cfg.vocab_size, (hparams["batch"], hparams["seq"]), generator=g).to(dev)
        tgt = torch.randint(0, cfg.vocab_size, (hparams["batch"], hparams["seq"]), generator=g).to(dev)
        opt.zero_grad()
        loss = model(idx, tgt)
        loss.backward()
        opt.step()
        reporter.broadcast(loss.detach(), step)           # device tensor: no host sync
    return float(loss)


if __name__ == "__main__":
    from maggy_b200 import experiment
    from maggy_b200.config import TorchDistributedConfig
    from maggy_b200.models.gpt2 import GPT2

    cfg = TorchDistributedConfig(module=GPT2, backend="torch", zero_lvl=2,
                                 hparams={"model": "gpt2-small", "lr": 3e-4, "steps": 20, "batch": 8, "seq": 1024})
    print(experiment.lagom(train_fn, cfg))
