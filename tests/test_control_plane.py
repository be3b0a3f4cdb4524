"""Control plane on the host-memory rings: protocol state machine, native median-rule policy engine,
reporter semantics (reference: maggy/core/rpc.py, maggy/core/reporter.py)."""
import os
import random
import threading
import time

import pytest

from maggy_b200 import Trial
from maggy_b200.core import exceptions, rpc
from maggy_b200.core.reporter import Reporter
from maggy_b200.earlystop import MedianStoppingRule


@pytest.fixture
def chan():
    d = rpc.DriverChannel(3, ring_slots=64, payload_bytes=1 << 16)
    yield d
    d.close()


def _drain(d, n, timeout=2.0):
    out, t0 = [], time.time()
    while len(out) < n and time.time() - t0 < timeout:
        out += d.next(0.05)
    return out


def test_register_assign_metric_final_roundtrip(chan):
    w = rpc.WorkerChannel(chan.name, 1)
    w.register()
    (reg,) = _drain(chan, 1)
    assert reg["type"] == "REG" and reg["partition_id"] == 1 and reg["aux"] == 0
    assert chan.worker_state(1) == rpc.W_REGISTERED and chan.worker_pid(1) == os.getpid()
    assert w.wait_assign(0.01) is None                       # nothing assigned yet
    tid = Trial({"lr": 0.1}).trial_id
    assert chan.assign(1, tid, {"params": {"lr": 0.1, "fn": lambda x: x * 2}})
    assert chan.assign(1, tid, {"params": {}}) is False      # slot still occupied
    got_tid, payload = w.wait_assign(1.0)
    assert got_tid == tid and payload["params"]["fn"](21) == 42
    for step, v in enumerate([0.1, 0.2, 0.3]):
        w.post_metric(tid, v, step)
    w.post_log(tid, "hello from the worker")
    msgs = _drain(chan, 4)        # one blob is in flight per worker: the driver acks it by reading it
    w.post_final(tid, 0.3, {"logs": "bye"})
    msgs += _drain(chan, 1)
    assert [m["type"] for m in msgs] == ["METRIC"] * 3 + ["LOG", "FINAL"]
    assert [m["value"] for m in msgs[:3]] == [0.1, 0.2, 0.3] and [m["step"] for m in msgs[:3]] == [0, 1, 2]
    assert msgs[3]["data"] == "hello from the worker"
    assert msgs[4]["data"] == {"metric": 0.3, "logs": "bye"} and msgs[4]["trial_id"] == tid
    assert all(m["t_ns"] > 0 for m in msgs)
    w.close()


def test_stop_word_and_global_stop(chan):
    w = rpc.WorkerChannel(chan.name, 0)
    w.register()
    tid, other = "00000000000000aa", "00000000000000bb"
    chan.assign(0, tid, {"params": {}})
    w.wait_assign(1.0)
    assert not w.check_stop(tid)
    chan.stop_trial(0, other)
    assert not w.check_stop(tid) and w.check_stop(other)
    chan.stop_trial(0, tid)
    assert w.check_stop(tid)
    chan.assign(0, other, {"params": {}})                     # a new assignment clears the stop word
    w.wait_assign(1.0)
    assert not w.check_stop(other)
    t0 = time.time()
    res = {}
    th = threading.Thread(target=lambda: res.setdefault("r", w.wait_assign(5.0)))
    th.start()
    time.sleep(0.05)
    chan.gstop()
    th.join(2)
    assert res["r"] is rpc.WorkerChannel.GSTOP and time.time() - t0 < 1.0   # futex wake, not a poll period
    w.close()


def test_reregistration_after_death_resets_slot(chan):
    w = rpc.WorkerChannel(chan.name, 2)
    w.register()
    chan.assign(2, "00000000000000cc", {"params": {"a": 1}})
    w.wait_assign(1.0)
    w.close()                                                 # "dies" while the trial is assigned
    w2 = rpc.WorkerChannel(chan.name, 2, attempt=1)
    w2.register()
    regs = [m for m in _drain(chan, 2) if m["type"] == "REG"]
    assert regs[-1]["aux"] == 1                               # driver sees attempt 1 => reschedule (BLACK)
    assert w2.wait_assign(0.01) is None                       # the consumed assignment is not replayed
    assert chan.assign(2, "00000000000000cc", {"params": {"a": 1}})
    assert w2.wait_assign(1.0)[0] == "00000000000000cc"
    w2.close()


def test_ring_wraparound_and_backpressure():
    d = rpc.DriverChannel(1, ring_slots=8, pump=False)
    w = rpc.WorkerChannel(d.name, 0)
    got = []
    stop = threading.Event()

    def consumer():
        while not stop.is_set() or len(got) < 200:
            got.extend(m["step"] for m in d.next(0.001) if m["type"] == "METRIC")
            if len(got) >= 200:
                break

    th = threading.Thread(target=consumer)
    th.start()
    for i in range(200):                                       # 25 laps around an 8-slot ring
        w.post_metric("00000000000000dd", float(i), i)
    stop.set()
    th.join(5)
    assert got == list(range(200))
    w.close()
    d.close()


def test_native_policy_equals_python_median_rule():
    """Same stream of trials through the native engine and the Python oracle: identical stop decisions."""
    rng = random.Random(7)
    for direction in ("max", "min"):
        d = rpc.DriverChannel(1, ring_slots=256, pump=False)
        d.configure_policy(True, direction, es_min=2, es_interval=1)
        w = rpc.WorkerChannel(d.name, 0)
        finals, native, oracle = [], [], []
        for n in range(25):
            t = Trial({"n": n})
            hist = [rng.random() for _ in range(rng.randint(1, 6))]
            stopped_native = False
            for step, v in enumerate(hist):
                w.post_metric(t.trial_id, v, step)
                msgs = [m for m in d.next(0) if m["type"] == "METRIC"]
                t.append_metric({"value": v, "step": step})
                want = (len(finals) > 2 and
                        MedianStoppingRule.earlystop_check(t, finals, direction) is not None)
                if msgs[0]["stopped"]:
                    stopped_native = True
                    assert w.check_stop(t.trial_id)
                if want or stopped_native:
                    oracle.append((n, step, want))
                    native.append((n, step, stopped_native))
                    break
            w.post_final(t.trial_id, hist[-1])
            d.next(0)
            t.final_metric = hist[-1]
            finals.append(t)
        assert native == oracle and any(s for _, _, s in native)
        assert d.policy_stops() == sum(1 for _, _, s in native if s)
        w.close()
        d.close()


def test_reporter_semantics(tmp_path, chan):
    w = rpc.WorkerChannel(chan.name, 0)
    w.register()
    rep = Reporter(str(tmp_path / "exec.log"), 0, 0, None, w)
    tid = "00000000000000ee"
    chan.assign(0, tid, {"params": {}})
    w.wait_assign(1.0)
    rep.set_trial_id(tid)
    rep.init_logger(str(tmp_path / "trial.log"))
    rep.broadcast(0.5)            # step defaults to previous + 1
    rep.broadcast(0.6)
    assert rep.step == 1 and rep.get_data()[0] == 0.6
    rep.broadcast(0.7, 5)
    with pytest.raises(exceptions.BroadcastStepValueError):
        rep.broadcast(0.8, 3)     # steps must not go backwards
    with pytest.raises(exceptions.BroadcastMetricTypeError):
        rep.broadcast("0.9")
    with pytest.raises(exceptions.BroadcastStepTypeError):
        rep.broadcast(0.9, "7")
    rep.log("a line for the driver", jupyter=True)
    msgs = _drain(chan, 3)
    chan.stop_trial(0, tid)
    with pytest.raises(exceptions.EarlyStopException) as e:
        rep.broadcast(0.9, 6)     # the very next broadcast raises
    assert e.value.metric == 0.9
    rep.reset()
    assert rep.step == -1 and rep.metric is None and rep.stop is False
    msgs += _drain(chan, 3)
    metrics = [(m["step"], m["value"]) for m in msgs if m["type"] == "METRIC"]
    assert metrics == [(0, 0.5), (1, 0.6), (5, 0.7), (6, 0.9)]
    assert any(m["type"] == "LOG" and "a line for the driver" in m["data"] for m in msgs)
    assert "a line for the driver" in (tmp_path / "trial.log").read_text()
    rep.close_logger()
    w.close()


def test_reporter_without_channel_and_early_stop_needs_a_metric(tmp_path):
    rep = Reporter(str(tmp_path / "e.log"), 0, 0, None, None)
    rep.early_stop()
    rep.broadcast(1.0)            # no metric had been reported before early_stop(): nothing happens
    rep.early_stop()
    with pytest.raises(exceptions.EarlyStopException):
        rep.broadcast(2.0)
    rep.close_logger()


def test_stop_latency_is_sub_millisecond(chan):
    """METRIC record -> native policy -> STOP word, measured with the ring's own monotonic stamps."""
    chan.configure_policy(True, "max", es_min=0, es_interval=1)
    w = rpc.WorkerChannel(chan.name, 0)
    good = Trial({"q": 1})
    for s in range(3):
        w.post_metric(good.trial_id, 0.9, s)
    w.post_final(good.trial_id, 0.9)
    _drain(chan, 4)
    lat = []
    for n in range(20):
        bad = Trial({"bad": n})
        w.post_metric(bad.trial_id, 0.1, 0)
        t0 = time.perf_counter()
        while not w.check_stop(bad.trial_id):
            assert time.perf_counter() - t0 < 1.0
        (m,) = [m for m in _drain(chan, 1) if m["type"] == "METRIC"]
        assert m["stopped"]
        lat.append((chan.stop_time_ns(0) - m["t_ns"]) / 1e6)
        w.post_final(bad.trial_id, 0.1)
        _drain(chan, 1)
    lat.sort()
    assert lat[len(lat) // 2] < 5.0, lat   # ms; the reference needs 1-2 heartbeat periods (1-2 s)
    w.close()



def test_prefetched_trial_is_parked_then_published():
    """mgr_stage parks the worker's next trial; mgr_publish_staged (what K9 does from the GPU on the worker's FINAL)
    hands it over without touching the payload; a direct assignment invalidates a parked entry."""
    from maggy_b200.core import rpc

    ch = rpc.DriverChannel(1, ring_slots=64, payload_bytes=1 << 16)
    try:
        w = rpc.WorkerChannel(ch.name, 0)
        w.register()
        assert any(m["type"] == "REG" for m in ch.next(0.5))
        assert ch.assign(0, "00000000000000a1", {"params": {"lr": 0.1}}, tag=1)
        assert ch.staged_state(0)[0] is None
        assert ch.stage(0, "00000000000000b2", {"params": {"lr": 0.2}}, tag=1)
        assert not ch.stage(0, "00000000000000c3", {"params": {"lr": 0.3}})          # slot taken
        assert ch.staged_state(0) == ("parked", "00000000000000b2")
        assert not ch.publish_staged(0)                                               # a1 not picked up yet
        tid, payload = w.wait_assign(1.0)
        assert tid == "00000000000000a1" and payload["params"]["lr"] == 0.1
        w.post_final(tid, 0.5, {"metric": 0.5})
        assert ch.publish_staged(0)
        assert ch.staged_state(0) == ("published", "00000000000000b2")
        tid, payload = w.wait_assign(1.0)
        assert tid == "00000000000000b2" and payload["params"]["lr"] == 0.2
        assert ch.reservations.get_assigned_trial(0) == "00000000000000b2"
        finals = [m for m in ch.next(0.2) if m["type"] == "FINAL"]
        assert finals and finals[0]["trial_id"] == "00000000000000a1"
        # park again, then assign directly: the parked entry is void and can be re-parked
        assert ch.stage(0, "00000000000000d4", {"params": {"lr": 0.4}})
        assert ch.assign(0, "00000000000000e5", {"params": {"lr": 0.5}})
        assert ch.staged_state(0)[0] is None and not ch.publish_staged(0)
        assert w.wait_assign(1.0)[0] == "00000000000000e5"
        assert ch.stage(0, "00000000000000d4", {"params": {"lr": 0.4}})
        w.close()
    finally:
        ch.close()


def test_executor_factory_runs_the_trial_loop_in_process(tmp_path, monkeypatch):
    """trial_executor_fn(...) (the reference's factory signature) -> executor_fn(partition_id): attaches to the control
    segment by name and serves trials; driven here by a bare DriverChannel (REG -> assign -> METRIC/FINAL -> GSTOP)."""
    import threading

    from maggy_b200.core import rpc
    from maggy_b200.core.environment.singleton import EnvSing
    from maggy_b200.core.executors.trial_executor import trial_executor_fn

    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "experiment_log"))
    EnvSing.reset()
    log_dir = str(tmp_path / "run")
    os.makedirs(log_dir, exist_ok=True)

    def train_fn(hparams, reporter):
        reporter.broadcast(hparams["x"] * 2.0, 0)
        return hparams["x"] * 2.0

    class Cfg:
        name, description, hb_interval, optimization_key, model, dataset = "f", "", 1, "Metric", None, None

    ch = rpc.DriverChannel(1, ring_slots=64, payload_bytes=1 << 16)
    try:
        fn = trial_executor_fn(train_fn, Cfg(), "optimization", "application_1_0001", 1, ch.name, 1, None, "Metric", log_dir)
        rc = {}
        t = threading.Thread(target=lambda: rc.setdefault("rc", fn(0)), daemon=True)
        t.start()
        seen, deadline = [], time.time() + 30
        assigned = False
        while time.time() < deadline and not any(m["type"] == "FINAL" for m in seen):
            for m in ch.next(0.05):
                seen.append(m)
                if m["type"] == "REG" and not assigned:
                    assert ch.assign(0, "00000000000000f1", {"params": {"x": 1.5}, "trial_type": "optimization"})
                    assigned = True
        finals = [m for m in seen if m["type"] == "FINAL"]
        assert finals and finals[0]["trial_id"] == "00000000000000f1" and finals[0]["data"]["metric"] == 3.0
        ch.gstop()
        t.join(10)
        assert not t.is_alive() and rc.get("rc") == 0
    finally:
        ch.close()
        EnvSing.reset()


def test_reference_server_and_client_method_names():
    """DriverChannel / WorkerChannel answer to the reference's Server / Client vocabulary (start, await_reservations,
    get_suggestion, finalize_metric, stop, MessageSocket.send / receive, get_assigned_trial_id)."""
    from maggy_b200.core import rpc

    server = rpc.OptimizationServer(1, ring_slots=64, payload_bytes=1 << 16)
    try:
        addr, n = server.start()
        client = rpc.Client(addr, 0)
        client.register()
        assert any(m["type"] == "REG" for m in rpc.MessageSocket.receive(server, 0.5))
        assert server.await_reservations(timeout=1)[0]["partition_id"] == 0 if "partition_id" in server.reservations.get()[0] \
            else server.reservations.done()
        assert client.await_reservations() and client.get_message("QUERY") is True
        stop = client.start_heartbeat(None, interval_s=0.01)
        assert server.assign(0, "00000000000000aa", {"params": {"lr": 0.5}})
        trial_id, params = client.get_suggestion()
        assert trial_id == "00000000000000aa" and params == {"lr": 0.5}
        assert server.get_assigned_trial_id(0) == "00000000000000aa"
        rpc.MessageSocket.send(client, {"type": "METRIC", "trial_id": trial_id, "data": {"value": 0.25, "step": 3}})
        client.finalize_metric(0.75)
        msgs = []
        for _ in range(20):
            msgs += server.next(0.05)
            if any(m["type"] == "FINAL" for m in msgs):
                break
        assert any(m["type"] == "METRIC" and m["value"] == 0.25 and m["step"] == 3 for m in msgs)
        assert any(m["type"] == "FINAL" and m["trial_id"] == trial_id and m["data"]["metric"] == 0.75 for m in msgs)
        with pytest.raises(NotImplementedError):
            client.get_message("TF_CONFIG")
        server.gstop()
        assert client.get_suggestion(timeout_s=2) == (None, None)
        client.stop()
        assert stop.is_set()
    finally:
        server.stop()
