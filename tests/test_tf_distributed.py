"""TensorFlow distributed stack (reference B4 / R4 / X3 / PT4): TF_CONFIG derivation, the strategy-scope model
wrapper, data sharding, and ``lagom(TfDistributedConfig)`` end to end on CPU workers with a TensorFlow stand-in."""
import json
import os
import sys

import pytest

from maggy_b200 import experiment
from maggy_b200.config import TfDistributedConfig
from maggy_b200.core.drivers.tf_distributed_training_driver import tf_config_for
from maggy_b200.core.environment.singleton import EnvSing
from maggy_b200.core.executors.tf_dist_executor import shard_data
from maggy_b200.core.reservations import MirroredReservations
from maggy_b200.core.tf_patching.tf_modules import get_wrapped_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cluster_spec_and_tf_config_roles():
    res = MirroredReservations(4, evaluator=True)
    for pid, port in [(2, 7002), (0, 7000), (3, 7003), (1, 7001)]:   # registration order decides the roles
        res.add({"partition_id": pid, "task_attempt": 0, "host_port": f"127.0.0.1:{port}", "trial_id": None})
    assert res.done()
    spec = res.get_spec()
    assert spec == {"chief": ["127.0.0.1:7002"], "evaluator": ["127.0.0.1:7001"],
                    "worker": ["127.0.0.1:7000", "127.0.0.1:7003"]}
    assert tf_config_for(spec, "127.0.0.1:7002")["task"] == {"type": "chief", "index": 0}
    assert tf_config_for(spec, "127.0.0.1:7003")["task"] == {"type": "worker", "index": 1}
    cfg = tf_config_for(spec, "127.0.0.1:7001")
    assert cfg["task"] == {"type": "evaluator", "index": 0} and cfg["cluster"] == spec
    with pytest.raises(KeyError):
        tf_config_for(spec, "127.0.0.1:1")
    two = MirroredReservations(2, evaluator=True)   # reference: with two executors the second one evaluates
    two.add({"partition_id": 0, "host_port": "a:1", "task_attempt": 0, "trial_id": None})
    two.add({"partition_id": 1, "host_port": "b:2", "task_attempt": 0, "trial_id": None})
    assert two.get_spec() == {"chief": ["a:1"], "evaluator": ["b:2"]}


def test_wrapped_model_is_built_under_the_strategy_scope():
    import contextlib

    class Strategy:
        depth = 0

        @contextlib.contextmanager
        def scope(self):
            Strategy.depth += 1
            yield
            Strategy.depth -= 1

    class Net:
        def __init__(self, units, act="relu"):
            self.units, self.act, self.depth_at_init = units, act, Strategy.depth

    Wrapped = get_wrapped_model(Net, Strategy(), is_chief=True)
    m = Wrapped(8, act="tanh")
    assert isinstance(m, Net) and (m.units, m.act) == (8, "tanh")
    assert m.depth_at_init == 1 and Strategy.depth == 0
    assert Wrapped._maggy_is_chief and Wrapped.__name__ == "MaggyNet"
    assert get_wrapped_model(None, Strategy(), False) is None
    with pytest.raises(TypeError):
        get_wrapped_model(Net(1), Strategy(), False)


def test_shard_data():
    assert shard_data(list(range(10)), 3, 1) == [1, 4, 7]
    xs, ys = shard_data((list(range(6)), list("abcdef")), 2, 0)
    assert xs == [0, 2, 4] and ys == ["a", "c", "e"]

    class DS:
        def shard(self, n, i):
            return ("sharded", n, i)

    assert shard_data(DS(), 4, 2) == ("sharded", 4, 2)
    assert shard_data(None, 4, 2) is None and shard_data([1, 2], 1, 0) == [1, 2]


class TinyModel:
    def __init__(self, width):
        self.width = width


def train_tf(model, dataset, hparams, reporter):
    import json as _json
    import os as _os

    role = _json.loads(_os.environ["TF_CONFIG"])["task"]["type"]
    net = model(hparams["width"])                       # built under strategy.scope() by the wrapper
    assert type(net).__name__ == "MaggyTinyModel" and net.width == 4
    for step in range(3):
        reporter.broadcast(float(len(dataset)) + step, step)
    reporter.log(f"role={role} shard={dataset}", True)
    return {"loss": float(sum(dataset)), "acc": 0.5}


@pytest.fixture()
def _tf_env(tmp_path, monkeypatch):
    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "experiment_log"))
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "3")
    monkeypatch.setenv("PYTHONPATH", os.pathsep.join([ROOT, os.path.join(ROOT, "tests", "fakes"),
                                                      os.environ.get("PYTHONPATH", "")]))
    EnvSing.reset()
    yield tmp_path
    EnvSing.reset()


def test_lagom_tf_distributed_end_to_end_with_stand_in(_tf_env, monkeypatch):
    monkeypatch.setenv("MAGGY_B200_TF_MODULE", "fake_tf")
    cfg = TfDistributedConfig(model=TinyModel, dataset=list(range(12)), process_data=lambda d: [x * 2 for x in d],
                              hparams={"width": 4}, mixed_precision=True, name="tf_e2e")
    res = experiment.lagom(train_tf, cfg)
    # chief and the single worker each train on half of the processed data, the evaluator sees all of it
    full = sum(x * 2 for x in range(12))
    expected = sorted([sum([x * 2 for x in range(12)][0::2]), sum([x * 2 for x in range(12)][1::2]), full])
    assert res["test result"] == pytest.approx(sum(expected) / 3)
    root = str(_tf_env / "experiment_log")
    app = os.path.join(root, sorted(os.listdir(root))[-1])
    run = os.path.join(app, sorted(os.listdir(app), key=int)[-1])   # RUN_ID keeps counting within a process
    meta = json.load(open(os.path.join(run, "maggy.json")))
    assert meta["experiment_type"] == "tf_dist" and meta["executors"] == 3
    assert sorted(meta["cluster"]) == ["chief", "evaluator", "worker"]
    logs = "".join(open(os.path.join(run, f"training_logs_{w}", "output.log")).read() for w in range(3))
    assert "role=chief" in logs and "role=worker" in logs and "role=evaluator" in logs


def test_lagom_tf_distributed_without_tensorflow_fails_with_a_pointer(_tf_env, monkeypatch):
    monkeypatch.setenv("MAGGY_B200_TF_MODULE", "definitely_not_tensorflow")
    cfg = TfDistributedConfig(model=TinyModel, dataset=[1, 2, 3], hparams={"width": 4})
    with pytest.raises(Exception) as e:
        experiment.lagom(train_tf, cfg)
    assert "TorchDistributedConfig" in str(e.value) and "not importable" in str(e.value)
