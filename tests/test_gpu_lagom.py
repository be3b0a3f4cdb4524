"""``experiment.lagom`` with real GPU workers: trials train a tiny GPT-2 on the sm_100a kernels and report their
loss as a CUDA tensor (published from the GPU by the K8 kernel)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def train_gpt2_tiny(hparams, reporter):
    import torch

    from maggy_b200.models.gpt2 import GPT2, GPT2Config
    from maggy_b200.parallel.arena import FlatParams, FusedAdamW

    dev = torch.device("cuda", torch.cuda.current_device())
    cfg = GPT2Config.named("gpt2-tiny")
    torch.manual_seed(0)
    model = GPT2(cfg)
    flat = FlatParams(model, device=dev)
    opt = FusedAdamW(flat, lr=hparams["lr"], weight_decay=0.0)
    idx = torch.randint(0, cfg.vocab_size, (4, 64), device=dev)
    tgt = torch.randint(0, cfg.vocab_size, (4, 64), device=dev)
    loss = None
    for step in range(12):
        opt.zero_grad()
        loss = model(idx, tgt)
        loss.backward()
        opt.step()
        reporter.broadcast(-loss.detach(), step)      # stays on the device
    return -float(loss)


def test_hpo_on_gpu_workers(tmp_path, monkeypatch):
    from maggy_b200 import Searchspace, experiment
    from maggy_b200.config import HyperparameterOptConfig
    from maggy_b200.core.environment.singleton import EnvSing

    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "experiment_log"))
    EnvSing.reset()
    sp = Searchspace(lr=("DOUBLE", [1e-4, 5e-3]))
    cfg = HyperparameterOptConfig(num_trials=3, optimizer="randomsearch", searchspace=sp, direction="max",
                                  es_policy="none", name="gpu_hpo", seed=4)
    res = experiment.lagom(train_gpt2_tiny, cfg)
    EnvSing.reset()
    assert res["num_trials"] == 3 and res["best_val"] > -7.5     # ln(1000) = 6.9: it learned the fixed batch a bit
    run = os.path.join(str(tmp_path / "experiment_log"), os.listdir(str(tmp_path / "experiment_log"))[0], "1")
    trial = json.load(open(os.path.join(run, res["best_id"], "trial.json")))
    assert trial["step_history"] == list(range(12))               # every device-side report reached the driver
    assert trial["metric_history"][-1] == pytest.approx(res["best_val"], abs=0.05)
    assert trial["metric_history"][-1] > trial["metric_history"][0]
