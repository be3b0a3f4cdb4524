"""Host timeline tracing (MAGGY_TRACE=1): worker trial ranges + driver scheduling instants in Chrome trace format."""
import os

from maggy_b200 import Searchspace, experiment
from maggy_b200.config import HyperparameterOptConfig
from maggy_b200.core.environment.singleton import EnvSing
from maggy_b200.utils import tracing


def train(hparams, reporter):
    from maggy_b200.utils import tracing as tr

    with tr.range("user region", lr=hparams["lr"]):
        reporter.broadcast(hparams["lr"], 0)
    return hparams["lr"]


def test_trace_files_are_written(tmp_path, monkeypatch):
    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "experiment_log"))
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "2")
    monkeypatch.setenv("MAGGY_TRACE", "1")
    EnvSing.reset()
    sp = Searchspace(lr=("DOUBLE", [0.0, 1.0]), units=("INTEGER", [8, 64]))
    res = experiment.lagom(train, HyperparameterOptConfig(num_trials=4, optimizer="randomsearch", searchspace=sp,
                                                          direction="max", es_policy="none", name="trace", seed=1))
    assert res["num_trials"] == 4
    env = EnvSing.get_instance()
    app = sorted(os.listdir(env.log_dir))[-1]
    run = os.path.join(env.log_dir, app, sorted(os.listdir(os.path.join(env.log_dir, app)), key=int)[-1])
    drv = tracing.load(os.path.join(run, "trace_driver.json"))
    assert sum(e["name"] == "assign" for e in drv) == 4 and sum(e["name"] == "final" for e in drv) == 4
    worker_events = []
    for f in os.listdir(run):
        if f.startswith("trace_worker"):
            worker_events += tracing.load(os.path.join(run, f))
    trials = [e for e in worker_events if e["name"].startswith("trial ")]
    assert len(trials) == 4 and all(e["ph"] == "X" and e["dur"] > 0 for e in trials)
    assert sum(e["name"] == "user region" for e in worker_events) == 4
    EnvSing.reset()


def test_tracing_is_free_when_off(monkeypatch):
    monkeypatch.delenv("MAGGY_TRACE", raising=False)
    with tracing.range("nothing"):
        pass
    tracing.instant("nothing")
    assert not tracing.enabled()
