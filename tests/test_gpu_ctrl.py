"""Device side of the control plane on a real GPU: K8 report_store (metric published from the GPU without a
host sync) and K9 persistent scheduler kernel (median rule evaluated on the device)."""
import ctypes
import time

import pytest
import torch

from maggy_b200 import Trial
from maggy_b200.core import rpc
from maggy_b200.core.reporter import Reporter
from maggy_b200.core import exceptions

pytestmark = pytest.mark.gpu


def _drain(d, n, timeout=5.0):
    out, t0 = [], time.time()
    while len(out) < n and time.time() - t0 < timeout:
        out += d.next(0.01)
    return out


def test_report_store_publishes_device_metric():
    d = rpc.DriverChannel(2, ring_slots=64)
    w = rpc.WorkerChannel(d.name, 1)
    w.register()
    _drain(d, 1)
    tid = Trial({"a": 1}).trial_id
    vals = [0.25, 1.5, -3.0]
    for step, (v, dt) in enumerate(zip(vals, (torch.float32, torch.bfloat16, torch.float64))):
        w.post_metric_device(tid, torch.tensor(v, device="cuda", dtype=dt), step)
    torch.cuda.synchronize()
    msgs = [m for m in _drain(d, 3) if m["type"] == "METRIC"]
    assert [(m["step"], m["value"], m["trial_id"], m["aux"]) for m in msgs] == [
        (0, 0.25, tid, 1), (1, 1.5, tid, 1), (2, -3.0, tid, 1)]
    assert all(m["partition_id"] == 1 and m["t_ns"] > 0 for m in msgs)
    # reporter.broadcast with a CUDA tensor goes through the same kernel and still honours STOP
    rep = Reporter(None, 1, 0, None, w)
    rep.set_trial_id(tid)
    loss = torch.tensor(0.75, device="cuda")
    rep.broadcast(loss, 10)
    d.stop_trial(1, tid)
    with pytest.raises(exceptions.EarlyStopException) as e:
        rep.broadcast(loss * 2, 11)
    assert e.value.metric == pytest.approx(1.5)
    torch.cuda.synchronize()
    got = [(m["step"], m["value"]) for m in _drain(d, 2) if m["type"] == "METRIC"]
    assert got == [(10, 0.75), (11, 1.5)]
    w.close()
    d.close()


def test_report_store_ring_overflow_drops_instead_of_blocking():
    d = rpc.DriverChannel(1, ring_slots=8, pump=False)
    w = rpc.WorkerChannel(d.name, 0)
    t = torch.tensor(1.0, device="cuda")
    for i in range(30):                       # nobody drains: 8 fit, 22 are dropped, the stream never stalls
        w.post_metric_device("00000000000000aa", t, i)
    torch.cuda.synchronize()
    msgs = d.next(0)
    assert [m["step"] for m in msgs] == list(range(8))
    assert int(d.lib.mgr_dev_dropped(d.h, 0)) == 22
    w.close()
    d.close()


def test_persistent_scheduler_kernel_applies_median_rule():
    from maggy_b200.core.sched_kernel import DeviceScheduler

    d = rpc.DriverChannel(2, ring_slots=64, pump=False)
    d.configure_policy(False, "max", 0, 1)     # host policy off: only the GPU kernel may stop trials
    w0, w1 = rpc.WorkerChannel(d.name, 0), rpc.WorkerChannel(d.name, 1)
    # In production the scheduler kernel lives in the driver process, which does no other CUDA work. Here the
    # "worker" shares the process, so everything that allocates / maps (implicit device syncs would wait for
    # the persistent kernel forever) happens BEFORE the kernel starts.
    w1._device_state()
    bad_metric = torch.tensor(0.1, device="cuda")
    torch.cuda.synchronize()
    sched = DeviceScheduler(d, direction="max", es_min=0, es_interval=1, max_steps=32, cap_final=64)
    sched.start()
    try:   # a persistent kernel left running keeps the process from exiting: always stop it
        good = Trial({"g": 1})
        for s in range(4):
            w0.post_metric(good.trial_id, 0.9, s)
        w0.post_final(good.trial_id, 0.9)
        time.sleep(0.2)
        bad = Trial({"b": 1})
        t_post = time.perf_counter()
        w1.post_metric_device(bad.trial_id, bad_metric, 0)
        while not w1.check_stop(bad.trial_id):
            assert time.perf_counter() - t_post < 5.0, sched.stats()
        latency_ms = (time.perf_counter() - t_post) * 1e3
        ok = Trial({"ok": 1})
        w0.post_metric(ok.trial_id, 0.95, 0)
        time.sleep(0.2)
        assert not w0.check_stop(ok.trial_id)
    finally:
        stats = sched.stop()
    assert stats["stops"] == 1 and stats["metrics"] >= 6 and stats["finals"] == 1 and stats["polls"] > 10
    assert latency_ms < 100.0
    for c in (w0, w1):
        c.close()
    d.close()
