"""Environment flavours (reference E1): Hopsworks REST registration / xattr metadata against a local HTTP server,
Databricks cluster facts, and the external LOG endpoint of a running experiment."""
import http.server
import json
import os
import threading
import time

import pytest

from maggy_b200 import Searchspace, experiment
from maggy_b200.config import HyperparameterOptConfig
from maggy_b200.core.environment.singleton import EnvSing
from maggy_b200.core.logserver import LogServer, poll_logs, recv_msg, send_msg


class _Hopsworks(http.server.BaseHTTPRequestHandler):
    calls = []

    def _handle(self):
        n = int(self.headers.get("Content-Length", 0))
        body = self.rfile.read(n).decode() if n else ""
        _Hopsworks.calls.append((self.command, self.path, self.headers.get("Authorization"), body))
        self.send_response(200)
        self.end_headers()
        self.wfile.write(b"{}")

    do_POST = do_PUT = _handle

    def log_message(self, *a):
        pass


@pytest.fixture()
def hopsworks(tmp_path, monkeypatch):
    _Hopsworks.calls = []
    srv = http.server.ThreadingHTTPServer(("127.0.0.1", 0), _Hopsworks)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    monkeypatch.setenv("REST_ENDPOINT", f"http://127.0.0.1:{srv.server_address[1]}")
    monkeypatch.setenv("HOPSWORKS_PROJECT_ID", "119")
    monkeypatch.setenv("HOPSWORKS_PROJECT_NAME", "demo")
    monkeypatch.setenv("HOPSWORKS_API_KEY", "k3y")
    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "Experiments"))
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "2")
    EnvSing.reset()
    yield srv
    srv.shutdown()
    EnvSing.reset()


def train(hparams, reporter):
    import time

    for i in range(3):
        reporter.broadcast(hparams["x"], i)
        time.sleep(0.01)
    reporter.log("hello from the worker", True)
    return hparams["x"]


def test_hopsworks_env_registers_driver_and_attaches_metadata(hopsworks):
    from maggy_b200.core.environment.hopsworks import HopsworksEnv

    env = EnvSing.get_instance()
    assert isinstance(env, HopsworksEnv)
    cfg = HyperparameterOptConfig(num_trials=4, optimizer="randomsearch", direction="max", es_policy="none",
                                  searchspace=Searchspace(x=("DOUBLE", [0.0, 1.0])), name="hops", seed=1)
    res = experiment.lagom(train, cfg)
    assert res["num_trials"] == 4
    posts = [c for c in _Hopsworks.calls if c[0] == "POST"]
    assert len(posts) == 1 and posts[0][1] == "/hopsworks-api/api/maggy/drivers" and posts[0][2] == "ApiKey k3y"
    reg = json.loads(posts[0][3])
    assert set(reg) == {"hostIp", "port", "appId", "secret"} and reg["port"] > 0 and len(reg["secret"]) == 16
    puts = [c for c in _Hopsworks.calls if c[0] == "PUT"]
    assert puts and all(p[1].startswith("/hopsworks-api/api/project/119/experiments/") for p in puts)
    final = json.loads(puts[-1][3])
    assert puts[-1][1].endswith("?type=FULL_UPDATE") and final["state"] == "FINISHED"
    assert final["metric"] == pytest.approx(res["best_val"])
    run = os.path.join(env.log_dir, os.listdir(env.log_dir)[0], "1") if False else None  # noqa: F841
    runs = [os.path.join(dp, f) for dp, _, fs in os.walk(env.log_dir) for f in fs if f == ".summary.json"]
    assert runs, "Hopsworks environments write .summary.json"
    summary = json.load(open(runs[0]))
    assert len(summary["combinations"]) == 4 and "parameters" in summary["combinations"][0]


def test_hopsworks_registration_failure_only_logs(tmp_path, monkeypatch, capsys):
    monkeypatch.setenv("REST_ENDPOINT", "http://127.0.0.1:9")   # nothing listens there
    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "Experiments"))
    EnvSing.reset()
    env = EnvSing.get_instance()

    class D:
        lines = []

        def log(self, m):
            self.lines.append(str(m))

    assert env.announce_driver(1234, "s" * 16, "application_1_0001", D()) is False
    assert any("No logging" in line or "No connection" in line for line in D.lines)
    EnvSing.reset()


def test_databricks_cluster_facts(tmp_path, monkeypatch):
    monkeypatch.setenv("DATABRICKS_ROOT_CONDA_ENV", "x")
    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "dbfs" / "maggy_log"))
    monkeypatch.delenv("REST_ENDPOINT", raising=False)
    EnvSing.reset()
    env = EnvSing.get_instance()
    from maggy_b200.core.environment.databricks import DatabricksEnv

    assert isinstance(env, DatabricksEnv) and env.project_path() == "/dbfs/"
    tag = "spark.databricks.clusterUsageTags."
    assert env.cluster_workers({tag + "clusterScalingType": "autoscaling", tag + "clusterMaxWorkers": "8"}) == 8
    assert env.cluster_workers({tag + "clusterWorkers": "4"}) == 4
    with pytest.raises(KeyError):
        env.cluster_workers({tag + "clusterScalingType": "autoscaling"})
    with pytest.raises(KeyError):
        env.cluster_workers({})
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "6")
    monkeypatch.setenv("DATABRICKS_CLUSTER_WORKERS", "2")
    assert env.get_executors() == 2
    EnvSing.reset()


def test_log_endpoint_protocol():
    class D:
        num_trials = 9
        result = {"num_trials": 3, "early_stopped": 1, "best_val": 0.75}

        def get_logs(self):
            return self.result, "line\\n"

    srv = LogServer(D(), secret="a" * 16).start()
    try:
        rep = poll_logs(srv.address, "a" * 16)
        assert rep == {"type": "OK", "ex_logs": "line\\n", "num_trials": 9, "to_date": 3, "stopped": 1, "metric": 0.75}
        import socket

        with socket.create_connection(srv.address, timeout=2) as s:        # unknown request type -> ERR
            send_msg(s, {"type": "NOPE", "secret": "a" * 16})
            assert recv_msg(s) == {"type": "ERR"}
        with socket.create_connection(srv.address, timeout=2) as s:        # wrong secret: dropped, no reply
            send_msg(s, {"type": "LOG", "secret": "b" * 16})
            s.settimeout(1.0)
            assert s.recv(16) == b""
    finally:
        srv.stop()


def test_log_endpoint_serves_a_running_experiment(tmp_path, monkeypatch):
    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "experiment_log"))
    monkeypatch.setenv("MAGGY_NUM_EXECUTORS", "2")
    monkeypatch.setenv("MAGGY_LOG_SERVER", "1")
    monkeypatch.delenv("REST_ENDPOINT", raising=False)
    EnvSing.reset()
    seen = {}

    def poller():
        # the endpoint's address and secret are in maggy.log as soon as the driver is up
        deadline = time.time() + 20
        while time.time() < deadline and "reply" not in seen:
            for dp, _, fs in os.walk(str(tmp_path / "experiment_log")):
                if "maggy.log" in fs:
                    text = open(os.path.join(dp, "maggy.log")).read()
                    if "LOG endpoint listening on" in text and "secret" in text:
                        port = int(text.split("LOG endpoint listening on ")[1].split()[0].rsplit(":", 1)[1])
                        secret = text.split("secret ")[1].split()[0]
                        try:
                            seen["reply"] = poll_logs(("127.0.0.1", port), secret)
                        except OSError:
                            pass
            time.sleep(0.05)

    t = threading.Thread(target=poller, daemon=True)
    t.start()

    def slow(hparams, reporter):
        import time as _t

        for i in range(40):
            reporter.broadcast(hparams["x"], i)
            _t.sleep(0.01)
        return hparams["x"]

    cfg = HyperparameterOptConfig(num_trials=6, optimizer="randomsearch", direction="max", es_policy="none",
                                  searchspace=Searchspace(x=("DOUBLE", [0.0, 1.0])), name="logsrv", seed=2)
    res = experiment.lagom(slow, cfg)
    t.join(timeout=5)
    EnvSing.reset()
    assert res["num_trials"] == 6
    assert seen.get("reply", {}).get("type") == "OK" and seen["reply"]["num_trials"] == 6


def test_reference_util_and_environment_hooks_exist(monkeypatch):
    """Names of the reference's util / BaseEnv surface that only make sense with Spark or TCP still answer
    (maggy/util.py:63-76,246-253,339-347; environment/base.py:137-150,217-222)."""
    import socket

    from maggy_b200 import util
    from maggy_b200.core import rpc
    from maggy_b200.core.environment.base import BaseEnv

    monkeypatch.delenv("MAGGY_WORKER_ID", raising=False)
    assert util.get_partition_attempt_id() == (0, 0) and util.find_spark() is None
    monkeypatch.setenv("MAGGY_WORKER_ID", "3")
    monkeypatch.setenv("MAGGY_TASK_ATTEMPT", "2")
    assert util.get_partition_attempt_id() == (3, 2)
    old = util.APP_ID
    util.APP_ID = None
    util.set_app_id("application_1_0001")
    util.set_app_id("application_2_0001")
    assert util.APP_ID == "application_1_0001"
    util.APP_ID = old
    env = BaseEnv()
    s = socket.socket()
    s, (host, port) = env.connect_host(s, None, None)
    assert port > 0 and host
    s.close()
    ch = rpc.DriverChannel(1, ring_slots=64, payload_bytes=1 << 16)
    try:
        w = env.get_client(ch.name, 0, 1, "secret", None)
        w.register()
        assert any(m["type"] == "REG" for m in ch.next(0.5))
        w.close()
    finally:
        ch.close()
