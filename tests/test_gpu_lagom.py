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


def train_constant_quality_on_device(hparams, reporter):
    import time

    import torch

    q = torch.tensor(float(hparams["q"]), device="cuda")
    for step in range(80):
        reporter.broadcast(q, step)          # a CUDA tensor: published by K8 (mailbox + host-mapped ring)
        time.sleep(0.002)
    return float(hparams["q"])


def test_device_scheduler_stops_trials_during_a_real_lagom_run(tmp_path, monkeypatch):
    """K9 is started by HyperparameterOptDriver.init and is the ONLY native policy engine here (host engine off):
    the STOP words that end the bad trials early were written by the persistent kernel."""
    from maggy_b200 import Searchspace, experiment
    from maggy_b200.config import HyperparameterOptConfig
    from maggy_b200.core.environment.singleton import EnvSing

    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "experiment_log"))
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "1")
    monkeypatch.setenv("MAGGY_B200_DEVICE_SCHED", "1")
    monkeypatch.setenv("MAGGY_B200_HOST_POLICY", "0")
    EnvSing.reset()
    sp = Searchspace(q=("DOUBLE", [0.0, 1.0]))
    cfg = HyperparameterOptConfig(num_trials=10, optimizer="randomsearch", searchspace=sp, direction="max",
                                  es_policy="median", es_interval=1, es_min=2, name="gpu_k9", seed=3)
    res = experiment.lagom(train_constant_quality_on_device, cfg)
    EnvSing.reset()
    root = str(tmp_path / "experiment_log")
    app = os.path.join(root, sorted(os.listdir(root))[-1])
    run = os.path.join(app, sorted(os.listdir(app), key=int)[-1])   # RUN_ID keeps counting within a process
    meta = json.load(open(os.path.join(run, "maggy.json")))
    log = open(os.path.join(run, "maggy.log")).read()
    k9 = meta.get("device_scheduler")
    assert k9 is not None, log[-3000:]
    assert res["num_trials"] == 10
    assert k9["stops"] >= 1 and k9["metrics"] > 50 and k9["finals"] >= 3, (k9, log[-2000:])
    assert res["early_stopped"] >= 1, (k9, log[-2000:])
    stopped = [json.load(open(os.path.join(run, t, "trial.json"))) for t in os.listdir(run)
               if os.path.exists(os.path.join(run, t, "trial.json"))]
    short = [t for t in stopped if t["early_stop"]]
    lens = sorted(len(t["metric_history"]) for t in short)
    assert short and lens[0] < 80, (lens, k9)     # (a STOP that lands on the very last report still counts as one)
    # the records K9 judged came out of the device mailbox (GPU memory reached through CUDA IPC), not host memory
    assert k9["mailbox"] and k9["mailbox_records"] > 50, k9


def train_budgeted_quality_on_device(hparams, reporter):
    """Quality = q; ``budget`` units of 6 reports each, published from the GPU."""
    import torch

    q = torch.tensor(float(hparams["q"]), device="cuda")
    steps = 6 * int(hparams.get("budget", 1))
    for step in range(steps):
        reporter.broadcast(q + 1e-3 * step, step)
    return float(hparams["q"]) + 1e-3 * (steps - 1)


def test_asha_rung_decisions_are_taken_by_the_device_scheduler(tmp_path, monkeypatch):
    """ASHA through lagom with K9 running: every FINAL of a tagged trial is filed in the rung tables ON THE GPU and the
    promote / fresh-configuration decision the driver applies comes back from the kernel (no early-stop policy is
    configured, K9 runs for the scheduling decision alone)."""
    from maggy_b200 import Searchspace, experiment
    from maggy_b200.config import HyperparameterOptConfig
    from maggy_b200.core.environment.singleton import EnvSing
    from maggy_b200.optimizer import Asha

    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "experiment_log"))
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "1")
    monkeypatch.setenv("MAGGY_B200_DEVICE_SCHED", "1")
    EnvSing.reset()
    sp = Searchspace(q=("DOUBLE", [0.0, 1.0]))
    cfg = HyperparameterOptConfig(num_trials=16, optimizer=Asha(reduction_factor=2, resource_min=1, resource_max=4),
                                  searchspace=sp, direction="max", es_policy="none", name="gpu_k9_asha", seed=5)
    res = experiment.lagom(train_budgeted_quality_on_device, cfg)
    EnvSing.reset()
    root = str(tmp_path / "experiment_log")
    app = os.path.join(root, sorted(os.listdir(root))[-1])
    run = os.path.join(app, sorted(os.listdir(app), key=int)[-1])
    meta = json.load(open(os.path.join(run, "maggy.json")))
    log = open(os.path.join(run, "maggy.log")).read()
    k9 = meta.get("device_scheduler")
    assert k9 is not None and "asha_on_device=yes" in log, log[-3000:]
    trials = [json.load(open(os.path.join(run, t, "trial.json"))) for t in os.listdir(run)
              if os.path.exists(os.path.join(run, t, "trial.json"))]
    budgets = sorted(t["params"]["budget"] for t in trials)
    assert budgets[-1] == 4 and budgets.count(2) >= 2, budgets          # promotions up to the closing rung happened
    # one device decision per finished trial (the last FINAL closes the experiment before its decision is read)
    assert k9["asha_decisions"] >= len(trials) - 1, (k9, len(trials))
    assert k9["asha_decisions_applied"] >= len(trials) - 2 and k9["asha_decision_fallbacks"] <= 1, k9
    assert k9["stops"] == 0
    # prefetched hand-off: when the rung scan said "fresh configuration", K9 itself published the trial the driver had
    # parked for that worker (most decisions of a rung-0-heavy ASHA run), and the driver adopted exactly those
    assert k9["handoffs"] >= 3 and k9["prefetched_handoffs_adopted"] == k9["handoffs"], k9
    assert "final_to_decision_ms_median" in k9
    # the promoted configurations are the ones the rung tables rank on top: parent of every budget-2 trial is in the
    # better half of the budget-1 trials finished so far (checked loosely: better than the rung's worst)
    worst_b1 = min(t["final_metric"] for t in trials if t["params"]["budget"] == 1)
    for t in trials:
        if t["params"]["budget"] == 2:
            assert t["params"]["q"] > worst_b1 - 1e-2
    assert res["num_trials"] == len(trials)
