"""Reservations API parity (reference: maggy/core/rpc.py:45-202) and its wiring into the native control segment."""
import threading

import pytest

from maggy_b200.core.reservations import MirroredReservations, Reservations


def test_reservations_basic_and_thread_safety():
    r = Reservations(4)
    assert r.remaining() == 4 and not r.done()
    ths = [threading.Thread(target=r.add, args=({"partition_id": i, "host_port": f"h:{i}", "task_attempt": 0},))
           for i in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert r.done() and r.remaining() == 0
    r.assign_trial(2, "abcd")
    assert r.get_assigned_trial(2) == "abcd" and r.get_assigned_trial(1) is None
    assert r.get()[3]["num_executors"] == 4
    assert r.get_assigned_trial(99) is None


def test_mirrored_reservations_cluster_spec():
    r = MirroredReservations(4, evaluator=True)
    for i in range(4):
        r.add({"partition_id": i, "host_port": f"10.0.0.{i}:2222"})
    spec = r.get_spec()
    assert spec["chief"] == ["10.0.0.0:2222"]
    assert spec["evaluator"] == ["10.0.0.3:2222"]
    assert spec["worker"] == ["10.0.0.1:2222", "10.0.0.2:2222"]


def test_driver_channel_keeps_reservations_in_sync():
    from maggy_b200.core import rpc

    ch = rpc.DriverChannel(2, ring_slots=64, payload_bytes=1 << 16)
    try:
        w = rpc.WorkerChannel(ch.name, 1)
        w.register()
        msgs = []
        for _ in range(50):
            msgs += ch.next(0.05)
            if any(m["type"] == "REG" for m in msgs):
                break
        assert any(m["type"] == "REG" and m["partition_id"] == 1 for m in msgs)
        assert ch.reservations.remaining() == 1
        assert ch.reservations.get()[1]["host_port"].startswith("shm:")
        assert ch.assign(1, "00000000000000aa", {"lr": 0.1})
        assert ch.reservations.get_assigned_trial(1) == "00000000000000aa"
        w.close()
    finally:
        ch.close()


def test_tf_front_end_needs_tensorflow_only_at_run_time(monkeypatch):
    """The TF stack is implemented (tests/test_tf_distributed.py); TensorFlow itself is optional and its absence
    surfaces as NotSupportedError when a worker tries to import it."""
    from maggy_b200.core.exceptions import NotSupportedError
    from maggy_b200.core.executors.tf_dist_executor import import_tensorflow
    from maggy_b200.core.tf_patching import get_wrapped_model

    monkeypatch.setenv("MAGGY_B200_TF_MODULE", "definitely_not_tensorflow")
    with pytest.raises(NotSupportedError):
        import_tensorflow()
    assert get_wrapped_model(None, None, False) is None


def test_runtime_mode_flags():
    from maggy_b200.core import config

    config.initialize()
    assert config.is_spark_available() is False and config.MODE == "native-worker-pool"
