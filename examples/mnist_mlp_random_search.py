#!/usr/bin/env python
"""BASELINE.json config 1: ``experiment.lagom`` random search over a 2-hyperparameter MNIST-shaped MLP with two
CPU workers (the reference's "Spark local[2]" plumbing configuration, no GPU involved).

The image has no network, so the data is synthetic and MNIST-shaped (784 features, 10 classes, labels from a
fixed random teacher network so there is something to learn). Prints one JSON line with trials/hour.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def train_fn(hparams, reporter):
    import torch

    torch.manual_seed(0)
    torch.set_num_threads(2)
    teacher = torch.nn.Sequential(torch.nn.Linear(784, 64), torch.nn.Tanh(), torch.nn.Linear(64, 10))
    x = torch.randn(4096, 784)
    with torch.no_grad():
        y = teacher(x).argmax(-1)
    xv, yv = x[3584:], y[3584:]
    x, y = x[:3584], y[:3584]
    model = torch.nn.Sequential(torch.nn.Linear(784, hparams["units"]), torch.nn.ReLU(),
                                torch.nn.Linear(hparams["units"], 10))
    opt = torch.optim.SGD(model.parameters(), lr=hparams["lr"], momentum=0.9)
    acc = 0.0
    for epoch in range(int(hparams.get("epochs", 5))):
        for i in range(0, x.shape[0], 256):
            opt.zero_grad()
            loss = torch.nn.functional.cross_entropy(model(x[i:i + 256]), y[i:i + 256])
            loss.backward()
            opt.step()
        with torch.no_grad():
            acc = (model(xv).argmax(-1) == yv).float().mean().item()
        reporter.broadcast(acc, epoch)
    return {"accuracy": acc}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=12)
    ap.add_argument("--workers", type=int, default=2)
    args = ap.parse_args()
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    from maggy_b200 import Searchspace, experiment
    from maggy_b200.config import HyperparameterOptConfig

    sp = Searchspace(lr=("DOUBLE", [1e-3, 0.3]), units=("INTEGER", [16, 256]))
    cfg = HyperparameterOptConfig(num_trials=args.trials, optimizer="randomsearch", searchspace=sp, direction="max",
                                  optimization_key="accuracy", es_policy="median", es_min=4, es_interval=1,
                                  name="mnist_mlp_random_search", num_workers=args.workers, seed=0)
    t0 = time.time()
    res = experiment.lagom(train_fn, cfg)
    wall = time.time() - t0
    print(json.dumps({"metric": "trials/hour (random search, 2-param MNIST-shaped MLP, 2 CPU workers)",
                      "value": res["num_trials"] / wall * 3600.0, "unit": "trials/hour", "trials": res["num_trials"],
                      "early_stopped": res["early_stopped"], "best_accuracy": res["best_val"],
                      "best_config": res["best_config"], "wall_s": wall, "data": "synthetic MNIST-shaped"}))


if __name__ == "__main__":
    main()
