"""End-to-end ``experiment.lagom`` runs on CPU worker processes over the shared-memory control plane
(BASELINE.json config 1: "experiment.lagom random-search 2-param MNIST MLP, CPU (plumbing, no GPU)")."""
import json
import os
import signal
import time

import pytest

from maggy_b200 import Searchspace, experiment
from maggy_b200.ablation import AblationStudy
from maggy_b200.config import (AblationConfig, BaseConfig, HyperparameterOptConfig, OptimizationConfig,
                               TfDistributedConfig, TorchDistributedConfig)
from maggy_b200.core.environment.singleton import EnvSing
from maggy_b200.core.exceptions import NotSupportedError


@pytest.fixture(autouse=True)
def _isolated_logdir(tmp_path, monkeypatch):
    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "experiment_log"))
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "2")
    EnvSing.reset()
    yield
    EnvSing.reset()


def _run_dir():
    env = EnvSing.get_instance()
    app = sorted(os.listdir(env.log_dir))[-1]
    run = sorted(os.listdir(os.path.join(env.log_dir, app)), key=int)[-1]
    return os.path.join(env.log_dir, app, run)


def train_quadratic(hparams, reporter):
    import time

    acc = 0.0
    for i in range(4):
        acc = 1.0 - abs(hparams["lr"] - 0.3) - 0.1 / (i + 1)
        reporter.broadcast(acc, i)
        time.sleep(0.005)
    reporter.log("done", jupyter=True)
    return {"Metric": acc, "extra": hparams["units"]}


def test_random_search_end_to_end_and_artefacts():
    sp = Searchspace(lr=("DOUBLE", [0.0, 1.0]), units=("INTEGER", [8, 64]))
    cfg = HyperparameterOptConfig(num_trials=8, optimizer="randomsearch", searchspace=sp, direction="max",
                                  es_policy="none", name="rs", seed=3)
    res = experiment.lagom(train_quadratic, cfg)
    assert set(res) >= {"best_id", "best_val", "best_config", "worst_id", "worst_val", "worst_config", "avg",
                        "metric_list", "num_trials", "early_stopped", "num_epochs", "trial_id"}
    assert res["num_trials"] == 8 and res["early_stopped"] == 0 and len(res["metric_list"]) == 8
    assert res["best_val"] == max(res["metric_list"]) and res["worst_val"] == min(res["metric_list"])
    run = _run_dir()
    for f in ("maggy.log", "optimizer.log", "result.json", "maggy.json", "executor_0_0.log", "executor_1_0.log"):
        assert os.path.exists(os.path.join(run, f)), f
    assert json.load(open(os.path.join(run, "result.json")))["best_id"] == res["best_id"]
    meta = json.load(open(os.path.join(run, "maggy.json")))
    assert meta["status"] == "FINISHED" and meta["controller"] == "RandomSearch" and meta["executors"] == 2
    tdir = os.path.join(run, res["best_id"])
    for f in ("trial.json", ".hparams.json", ".outputs.json", ".metric", "output.log"):
        assert os.path.exists(os.path.join(tdir, f)), f
    trial = json.load(open(os.path.join(tdir, "trial.json")))
    assert trial["__class__"] == "Trial" and trial["status"] == "FINALIZED" and trial["trial_id"] == res["best_id"]
    assert trial["step_history"] == [0, 1, 2, 3] and len(trial["metric_history"]) == 4
    assert trial["final_metric"] == pytest.approx(res["best_val"])
    assert json.load(open(os.path.join(tdir, ".hparams.json"))) == res["best_config"]
    outs = json.load(open(os.path.join(tdir, ".outputs.json")))
    assert outs["Metric"] == pytest.approx(res["best_val"]) and "log" in outs and "extra" in outs
    assert float(open(os.path.join(tdir, ".metric")).read()) == pytest.approx(res["best_val"])
    assert "Maggy Optimization 8/8" in open(os.path.join(run, "maggy.log")).read()


def train_slow_when_bad(hparams, reporter):
    import time

    for i in range(200):
        reporter.broadcast(hparams["q"], i)   # constant quality: bad ones sit below the median forever
        time.sleep(0.002)
    return hparams["q"]


def test_median_early_stopping_cuts_bad_trials_short():
    sp = Searchspace(q=("DOUBLE", [0.0, 1.0]))
    cfg = OptimizationConfig(num_trials=14, optimizer="randomsearch", searchspace=sp, direction="max",
                             es_policy="median", es_interval=1, es_min=3, name="es", seed=11)
    t0 = time.time()
    res = experiment.lagom(train_slow_when_bad, cfg)
    assert res["num_trials"] == 14 and res["early_stopped"] >= 3
    run = _run_dir()
    stopped = []
    for tid in os.listdir(run):
        p = os.path.join(run, tid, "trial.json")
        if os.path.exists(p):
            t = json.load(open(p))
            if t["early_stop"]:
                stopped.append(len(t["metric_history"]))
    # stopped within a few reports of becoming eligible (trials that started before `es_min` trials had
    # finished can only be judged late)
    assert len([n for n in stopped if n < 20]) >= 3, stopped
    assert time.time() - t0 < 60


class _StopBelowHalf(__import__("maggy_b200.earlystop", fromlist=["AbstractEarlyStop"]).AbstractEarlyStop):
    """Custom policy: stop every trial whose constant quality is below 0.5 (the reference accepts instances and
    classes alike: optimization_driver.py:657-691)."""

    @staticmethod
    def earlystop_check(to_check, finalized_trials, direction):
        return to_check.trial_id if to_check.metric_history and to_check.metric_history[-1] < 0.5 else None


@pytest.mark.parametrize("as_instance", [True, False])
def test_custom_early_stop_policy_fires(as_instance):
    sp = Searchspace(q=("DOUBLE", [0.0, 1.0]))
    policy = _StopBelowHalf() if as_instance else _StopBelowHalf
    cfg = OptimizationConfig(num_trials=10, optimizer="randomsearch", searchspace=sp, direction="max",
                             es_policy=policy, es_interval=1, es_min=1, name="es_custom", seed=5)
    res = experiment.lagom(train_slow_when_bad, cfg)
    assert res["num_trials"] == 10
    run = _run_dir()
    n_bad_late = n_stopped = 0
    for tid in os.listdir(run):
        pth = os.path.join(run, tid, "trial.json")
        if os.path.exists(pth):
            t = json.load(open(pth))
            n_stopped += bool(t["early_stop"])
            if t["early_stop"]:
                assert t["metric_history"][-1] < 0.5
    assert n_stopped >= 1 and res["early_stopped"] == n_stopped
    assert "takes no arguments" not in open(os.path.join(run, "maggy.log")).read()


def test_no_stopping_rule_is_silent():
    sp = Searchspace(q=("DOUBLE", [0.0, 1.0]))
    cfg = OptimizationConfig(num_trials=6, optimizer="randomsearch", searchspace=sp, direction="max",
                             es_policy="none", es_interval=1, es_min=1, name="es_none", seed=5)
    res = experiment.lagom(train_quadratic_q, cfg)
    assert res["early_stopped"] == 0
    log = open(os.path.join(_run_dir(), "maggy.log")).read()
    assert "takes no arguments" not in log and "Trials to stop" not in log


def train_quadratic_q(hparams, reporter):
    for i in range(6):
        reporter.broadcast(hparams["q"], i)
    return hparams["q"]


def test_asha_and_gridsearch_run():
    sp = Searchspace(lr=("DOUBLE", [0.0, 1.0]), units=("INTEGER", [8, 64]))
    from maggy_b200.optimizer import Asha

    def train(hparams, reporter):
        for i in range(int(hparams["budget"])):
            reporter.broadcast(1.0 - abs(hparams["lr"] - 0.3), i)
        return 1.0 - abs(hparams["lr"] - 0.3)

    res = experiment.lagom(train, HyperparameterOptConfig(num_trials=16, optimizer=Asha(2, 1, 4), searchspace=sp,
                                                          direction="max", es_policy="none", name="asha", seed=1))
    run = _run_dir()
    budgets = sorted(json.load(open(os.path.join(run, d, "trial.json")))["params"]["budget"]
                     for d in os.listdir(run) if os.path.exists(os.path.join(run, d, "trial.json")))
    assert budgets[0] == 1 and budgets[-1] == 4 and res["num_trials"] == len(budgets) <= 16
    gsp = Searchspace(a=("DISCRETE", [1, 2, 3]), b=("CATEGORICAL", ["x", "y"]))
    res = experiment.lagom(lambda hparams: float(hparams["a"]) + (0.5 if hparams["b"] == "y" else 0.0),
                           HyperparameterOptConfig(num_trials=1, optimizer="gridsearch", searchspace=gsp,
                                                   direction="max", es_policy="none", name="grid"))
    assert res["num_trials"] == 6 and res["best_val"] == 3.5 and res["best_config"] == {"a": 3, "b": "y"}


def test_base_config_runs_in_process():
    seen = {}

    def train(reporter):
        reporter.broadcast(0.25)
        seen["pid"] = os.getpid()
        return {"accuracy": 0.9, "loss": 0.1}

    res = experiment.lagom(train)  # no config -> BaseConfig
    assert res == {"test result": {"accuracy": [0.9], "loss": [0.1]}} and seen["pid"] == os.getpid()
    assert experiment.lagom(lambda: 3, BaseConfig(name="n")) == {"test result": 3.0}


def test_ablation_loco_end_to_end():
    import torch

    def base_model():
        return torch.nn.Sequential()  # replaced below; LOCO needs named modules

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.stem = torch.nn.Linear(4, 8)
            self.block_a = torch.nn.Linear(8, 8)
            self.block_b = torch.nn.Linear(8, 8)
            self.head = torch.nn.Linear(8, 1)

        def forward(self, x):
            return self.head(self.block_b(self.block_a(self.stem(x))))

    def dataset(ablated_feature=None):
        feats = [f for f in ("f0", "f1", "f2", "f3") if f != ablated_feature]
        return feats

    study = AblationStudy("ds", 1, label_name="y", dataset_generator=dataset)
    study.features.include("f1", ["f2"])
    study.model.layers.include("block_a")
    study.model.layers.include_groups(prefix="block")
    study.model.set_base_model_generator(Net)

    def train(hparams, reporter):
        import torch

        model = hparams["model_function"]()
        feats = hparams["dataset_function"]()
        n_identity = sum(isinstance(m, torch.nn.Identity) for m in model.modules())
        return float(len(feats)) + 0.1 * n_identity

    res = experiment.lagom(train, AblationConfig(ablation_study=study, ablator="loco", direction="max", name="abl"))
    assert res["num_trials"] == 5      # base + 2 features + 1 layer + 1 prefix group
    assert sorted(round(m, 1) for m in res["metric_list"]) == [3.0, 3.0, 4.0, 4.1, 4.2]
    assert res["best_config"] == {"ablated_feature": "None", "ablated_layer": "Layers prefixed block"}
    run = _run_dir()
    hp = json.load(open(os.path.join(run, res["best_id"], ".hparams.json")))
    assert hp == {"ablated_feature": "None", "ablated_layer": "Layers prefixed block"}


def train_until_killed(hparams, reporter):
    import os
    import time

    marker = hparams["marker"]
    if hparams["victim"] and not os.path.exists(marker):
        open(marker, "w").write(str(os.getpid()))
        time.sleep(30)             # the test kills this process while the trial is running
    for i in range(3):
        reporter.broadcast(float(hparams["v"]), i)
    return float(hparams["v"])


def test_worker_death_reschedules_the_trial(tmp_path):
    marker = str(tmp_path / "victim.pid")
    sp = Searchspace(v=("DISCRETE", [1, 2, 3, 4]), victim=("DISCRETE", [True]), marker=("CATEGORICAL", [marker]))
    cfg = HyperparameterOptConfig(num_trials=1, optimizer="gridsearch", searchspace=sp, direction="max",
                                  es_policy="none", name="kill")
    import threading

    def killer():
        t0 = time.time()
        while not os.path.exists(marker) and time.time() - t0 < 60:
            time.sleep(0.05)
        time.sleep(0.2)
        os.kill(int(open(marker).read()), signal.SIGKILL)

    th = threading.Thread(target=killer)
    th.start()
    res = experiment.lagom(train_until_killed, cfg)
    th.join()
    assert res["num_trials"] == 4 and sorted(res["metric_list"]) == [1.0, 2.0, 3.0, 4.0]
    log = open(os.path.join(_run_dir(), "maggy.log")).read()
    assert "died unexpectedly" in log and "rescheduling lost trial" in log


def test_train_fn_exception_fails_the_experiment():
    sp = Searchspace(x=("DOUBLE", [0.0, 1.0]))

    def train(hparams):
        raise ValueError("boom in user code")

    with pytest.raises(RuntimeError, match="boom in user code"):
        experiment.lagom(train, HyperparameterOptConfig(num_trials=2, optimizer="randomsearch", searchspace=sp,
                                                        es_policy="none", name="err"))


def test_config_validation():
    sp = Searchspace(x=("DOUBLE", [0.0, 1.0]))
    with pytest.raises(ValueError):
        HyperparameterOptConfig(num_trials=0, optimizer="randomsearch", searchspace=sp)
    with pytest.raises(ValueError):
        HyperparameterOptConfig(num_trials=2, optimizer="randomsearch", searchspace=sp, direction="up")
    with pytest.raises(Exception, match="Unknown Optimizer"):
        experiment.lagom(lambda: 1, HyperparameterOptConfig(num_trials=2, optimizer="sgd", searchspace=sp))
    with pytest.raises(ValueError):
        TorchDistributedConfig(module=None, backend="horovod")
    with pytest.raises(ValueError):
        TorchDistributedConfig(module=None, backend="deepspeed")
    # TensorFlow is optional: without it a TF experiment fails inside the workers, with a pointer to the torch config
    with pytest.raises(RuntimeError, match="TorchDistributedConfig"):
        experiment.lagom(lambda: 1, TfDistributedConfig())
    with pytest.raises(TypeError):
        experiment.lagom(lambda: 1, object())


def train_dist(module, hparams, reporter):
    import torch
    import torch.distributed as dist

    torch.manual_seed(0)
    model = module(hparams["width"])
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    rank, world = dist.get_rank(), dist.get_world_size()
    x = torch.full((4, hparams["width"]), float(rank + 1))
    for step in range(3):
        opt.zero_grad()
        loss = model(x).pow(2).mean()
        loss.backward()
        opt.step()
        reporter.broadcast(loss.item(), step)
    w = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    assert all(torch.allclose(g, gathered[0]) for g in gathered)   # DDP kept the replicas in sync
    return float(rank)


def test_torch_distributed_training_over_gloo():
    import torch

    class Tiny(torch.nn.Module):
        def __init__(self, width):
            super().__init__()
            self.l = torch.nn.Linear(width, 1)

        def forward(self, x):
            return self.l(x)

    cfg = TorchDistributedConfig(module=Tiny, hparams={"width": 8}, backend="torch", zero_lvl=1, name="td")
    res = experiment.lagom(train_dist, cfg)
    assert res == {"test result": 0.5}     # mean of the ranks' metrics (0 and 1)
    run = _run_dir()
    assert os.path.exists(os.path.join(run, "training_logs_0", "output.log"))
    assert os.path.exists(os.path.join(run, "training_logs_1", ".outputs.json"))


def train_gang(hparams, reporter):
    import torch
    import torch.distributed as dist

    assert dist.is_initialized() and dist.get_world_size() == 2
    t = torch.tensor([float(dist.get_rank() + 1)])
    dist.all_reduce(t)                       # both gang members run the same trial
    for i in range(2):
        reporter.broadcast(hparams["x"] + t.item(), i)
    return hparams["x"] + float(dist.get_rank())   # leader reports the gang mean


def test_multi_process_trial_gangs(monkeypatch):
    """gpus_per_trial=2: every trial runs on a 2-process gang with a ready process group (gloo on CPU)."""
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "4")
    sp = Searchspace(x=("DOUBLE", [0.0, 1.0]))
    cfg = HyperparameterOptConfig(num_trials=4, optimizer="randomsearch", searchspace=sp, direction="max",
                                  es_policy="none", name="gang", gpus_per_trial=2, seed=5)
    res = experiment.lagom(train_gang, cfg)
    assert res["num_trials"] == 4
    run = _run_dir()
    meta = json.load(open(os.path.join(run, "maggy.json")))
    assert meta["executors"] == 2          # 4 processes = 2 gangs of 2
    for tid in os.listdir(run):
        p = os.path.join(run, tid, "trial.json")
        if os.path.exists(p):
            t = json.load(open(p))
            assert t["final_metric"] == pytest.approx(t["params"]["x"] + 0.5)      # mean of x+0 and x+1
            assert t["metric_history"] == pytest.approx([t["params"]["x"] + 3.0] * 2)


def train_gang_early_stop(hparams, reporter):
    import torch
    import torch.distributed as dist

    for i in range(100):
        t = torch.ones(1)
        dist.all_reduce(t)                    # a member that left early would dead-lock the other one here
        reporter.broadcast(hparams["q"], i)
        import time

        time.sleep(0.003)
    return hparams["q"]


def test_gang_members_stop_together(monkeypatch):
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "2")
    sp = Searchspace(q=("DOUBLE", [0.0, 1.0]))
    cfg = HyperparameterOptConfig(num_trials=8, optimizer="randomsearch", searchspace=sp, direction="max",
                                  es_policy="median", es_interval=1, es_min=2, name="gang_es", gpus_per_trial=2, seed=2)
    res = experiment.lagom(train_gang_early_stop, cfg)
    assert res["num_trials"] == 8 and res["early_stopped"] >= 1


def test_gang_ablation_payload_with_closures(monkeypatch):
    """Ablation payloads carry local closures: the gang leader must be able to forward them to followers."""
    import torch

    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "2")

    def base():
        return torch.nn.Sequential(torch.nn.Linear(2, 2), torch.nn.ReLU(), torch.nn.Linear(2, 1))

    study = AblationStudy("ds", 1, label_name="y", dataset_generator=lambda: None)
    study.model.set_base_model_generator(base)
    study.model.layers.include("1")

    def train(hparams):
        import torch.distributed as dist

        model = hparams["model_function"]()
        return float(sum(isinstance(m, torch.nn.Identity) for m in model.modules())) + 0.0 * dist.get_world_size()

    res = experiment.lagom(train, AblationConfig(ablation_study=study, direction="max", name="gang_abl", gpus_per_trial=2))
    assert res["num_trials"] == 2 and sorted(res["metric_list"]) == [0.0, 1.0]


def train_counted(hparams, reporter):
    import os

    with open(os.environ["MAGGY_TEST_COUNTER"], "a") as f:
        f.write("x")
    acc = 1.0 - abs(hparams["lr"] - 0.3)
    reporter.broadcast(acc, 0)
    return acc


def test_experiment_resume_replays_finished_trials(tmp_path, monkeypatch):
    """Additive capability: re-running with ``resume_from=<old run dir>`` does not execute the
    configurations that already finished there."""
    counter = tmp_path / "count.txt"
    counter.write_text("")
    monkeypatch.setenv("MAGGY_TEST_COUNTER", str(counter))
    sp = Searchspace(lr=("DOUBLE", [0.0, 1.0]), units=("INTEGER", [8, 64]))

    def cfg(**kw):
        return HyperparameterOptConfig(num_trials=6, optimizer="randomsearch", searchspace=sp, direction="max",
                                       es_policy="none", name="resume", seed=11, **kw)

    first = experiment.lagom(train_counted, cfg())
    assert len(counter.read_text()) == 6
    old_run = _run_dir()
    # simulate a crash that lost two trials
    lost = [d for d in sorted(os.listdir(old_run)) if os.path.isfile(os.path.join(old_run, d, "trial.json"))][:2]
    for d in lost:
        os.remove(os.path.join(old_run, d, "trial.json"))
    second = experiment.lagom(train_counted, cfg(resume_from=old_run))
    assert len(counter.read_text()) == 6 + 2          # only the two lost configurations ran again
    assert second["num_trials"] == 6
    assert sorted(second["metric_list"]) == pytest.approx(sorted(first["metric_list"]))
    assert second["best_id"] == first["best_id"]


def train_with_poisoned_records(hparams, reporter):
    """Fault injection (SURVEY.md §5.3 "poison a ring slot"): garbage reaches the driver between valid reports."""
    ch = reporter.channel
    ch.lib.mgr_post(ch.h, ch.worker, 99, 0xDEADBEEF, 7, 123.0, 0)            # unknown record type
    ch.post_metric("00000000deadbeef", 1e9, 3)                               # METRIC for a trial nobody scheduled
    acc = 1.0 - abs(hparams["lr"] - 0.3)
    reporter.broadcast(acc, 0)
    ch.lib.mgr_post(ch.h, ch.worker, 3, 0xDEADBEEF, 0, 5.0, 0)               # FINAL for an unknown trial (no blob)
    return acc


def test_driver_survives_poisoned_records():
    sp = Searchspace(lr=("DOUBLE", [0.0, 1.0]), units=("INTEGER", [8, 64]))
    cfg = HyperparameterOptConfig(num_trials=4, optimizer="randomsearch", searchspace=sp, direction="max",
                                  es_policy="median", es_min=1, name="poison", seed=5)
    res = experiment.lagom(train_with_poisoned_records, cfg)
    assert res["num_trials"] == 4 and len(res["metric_list"]) == 4
    assert all(0.0 <= m <= 1.0 for m in res["metric_list"])                  # the 1e9 / 5.0 garbage never became a result


def test_asha_with_a_stand_in_for_the_device_scheduler(monkeypatch):
    """The driver side of the device-scheduler protocol, on CPU: a stand-in plays K9 with the SAME rung-table code
    (mg_asha.h through the host library) — it files every tagged FINAL, decides promote / fresh, and publishes the
    parked trial itself when the answer is "fresh" (mgr_publish_staged = what the kernel does with five stores). The
    driver must adopt the hand-offs, apply the promotions, keep parking prefetched trials and finish the experiment with
    the same artefacts as the host-decided run."""
    import ctypes

    from maggy_b200.core import rpc
    from maggy_b200.core.drivers.optimization_driver import HyperparameterOptDriver
    from maggy_b200.optimizer import Asha

    lib = rpc._host()
    state = {}

    def fake_start(self):      # instead of launching the CUDA kernel
        c = self.controller
        state["tab"] = lib.mgr_asha_create(c.reduction_factor, c.max_rung, 256, 1 if self.direction == "max" else -1)
        state["dec"], state["seq"] = {}, {}
        self._device_asha = True

    orig_final = HyperparameterOptDriver._final_msg_callback

    def final_with_k9(self, msg):   # "K9" sees the FINAL record before the driver's digest thread handles the message
        trial = self.get_trial(msg["trial_id"])
        if trial is not None and getattr(trial, "_device_tagged", False):
            w = msg["partition_id"]
            parent, rung = ctypes.c_uint64(), ctypes.c_uint32()
            metric = (msg["data"] or {}).get("metric", msg["value"])
            kind = lib.mgr_asha_final(state["tab"], self.controller.rung_of(trial), int(trial.trial_id, 16), float(metric),
                                      ctypes.byref(parent), ctypes.byref(rung))
            if kind == Asha.NEW and self.channel.staged_state(w)[0] == "parked" and self.channel.publish_staged(w):
                kind = Asha.NEW_PUBLISHED
            state["seq"][w] = state["seq"].get(w, 0) + 1
            state["dec"][(w, state["seq"][w])] = (kind, rung.value, f"{parent.value:016x}", msg["t_ns"] + 1000)
        return orig_final(self, msg)

    monkeypatch.setattr(HyperparameterOptDriver, "_start_device_scheduler", fake_start)
    monkeypatch.setattr(HyperparameterOptDriver, "_final_msg_callback", final_with_k9)
    monkeypatch.setattr(rpc.DriverChannel, "device_decision",
                        lambda self, worker, expected, timeout_s=0.002: state["dec"].get((worker, expected)))
    holder = {}
    orig_json = HyperparameterOptDriver.json

    def spy_json(self):
        holder["d"] = self
        return orig_json(self)

    monkeypatch.setattr(HyperparameterOptDriver, "json", spy_json)

    def train(hparams, reporter):
        for i in range(int(hparams["budget"])):
            reporter.broadcast(1.0 - abs(hparams["lr"] - 0.3), i)
        return 1.0 - abs(hparams["lr"] - 0.3)

    sp = Searchspace(lr=("DOUBLE", [0.0, 1.0]))
    res = experiment.lagom(train, HyperparameterOptConfig(num_trials=24, optimizer=Asha(2, 1, 4), searchspace=sp,
                                                          direction="max", es_policy="none", name="asha_k9", seed=2))
    d = holder["d"]
    run = _run_dir()
    trials = [json.load(open(os.path.join(run, t, "trial.json"))) for t in os.listdir(run)
              if os.path.exists(os.path.join(run, t, "trial.json"))]
    budgets = sorted(t["params"]["budget"] for t in trials)
    assert budgets[0] == 1 and budgets[-1] == 4 and budgets.count(2) >= 2 and res["num_trials"] == len(trials)
    assert d.device_decisions >= len(trials) - 2 and d.device_decision_fallbacks <= 1
    assert d.device_handoffs >= 2                     # "fresh" decisions were handed over by the stand-in
    handed = [t for t in trials if t["params"]["budget"] == 1]
    assert len(handed) >= d.device_handoffs
    # every promoted configuration carries its parent's hyperparameters and doubled budget
    by_lr = {}
    for t in trials:
        by_lr.setdefault(round(t["params"]["lr"], 12), []).append(t["params"]["budget"])
    assert any(sorted(b)[:2] == [1, 2] for b in by_lr.values())
    lib.mgr_asha_free(state["tab"])
