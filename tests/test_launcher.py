"""Native worker launcher (csrc/runtime/launcher.cpp): spawn with env / GPU pinning, exit codes, signals,
process-group kill — the per-GPU worker pool that replaces Spark executors."""
import ctypes
import os
import signal
import sys
import time

from maggy_b200.core import workerpool


def _spawn(lib, pool, slot, code, env=(), gpus="-1", log=""):
    argv = workerpool._blob([sys.executable, "-c", code])
    return lib.mgl_spawn(pool, slot, argv, workerpool._blob(list(env)), gpus.encode(), 0, 1, log.encode())


def test_spawn_env_exit_code_and_log(tmp_path):
    lib = workerpool._launcher()
    pool = lib.mgl_create(2)
    log = str(tmp_path / "w.log")
    code = "import os,sys; print(os.environ['FOO'], repr(os.environ.get('CUDA_VISIBLE_DEVICES'))); sys.exit(7)"
    pid = _spawn(lib, pool, 0, code, env=["FOO=bar"], gpus="-1", log=log)
    assert pid > 0 and lib.mgl_pid(pool, 0) == pid
    assert lib.mgl_wait(pool, 0, 10000) == 0
    assert lib.mgl_exit_code(pool, 0) == 7 and lib.mgl_term_signal(pool, 0) == 0
    assert open(log).read().strip() == "bar ''"          # "-1" hides every GPU
    pid = _spawn(lib, pool, 1, "import os; print(os.environ['CUDA_VISIBLE_DEVICES'])", gpus="3,5", log=log)
    assert lib.mgl_wait(pool, 1, 10000) == 0 and lib.mgl_exit_code(pool, 1) == 0
    assert open(log).read().strip().splitlines()[-1] == "3,5"
    assert lib.mgl_spawn_count(pool, 0) == 1 and lib.mgl_poll(pool, 0) == 0
    lib.mgl_destroy(pool)


def test_kill_takes_the_whole_process_group(tmp_path):
    lib = workerpool._launcher()
    pool = lib.mgl_create(1)
    marker = tmp_path / "child.pid"
    code = (f"import subprocess,sys,time; p=subprocess.Popen([sys.executable,'-c','import time; time.sleep(60)']);"
            f"open({str(marker)!r},'w').write(str(p.pid)); time.sleep(60)")
    _spawn(lib, pool, 0, code)
    t0 = time.time()
    while not marker.exists() and time.time() - t0 < 10:
        time.sleep(0.05)
    child = int(marker.read_text())
    assert lib.mgl_poll(pool, 0) == 1
    lib.mgl_kill(pool, 0, signal.SIGKILL)
    assert lib.mgl_wait(pool, 0, 5000) == 0
    assert lib.mgl_term_signal(pool, 0) == signal.SIGKILL and lib.mgl_exit_code(pool, 0) == 128 + signal.SIGKILL
    time.sleep(0.2)
    alive = True
    try:
        os.kill(child, 0)
        alive = open(f"/proc/{child}/stat").read().split()[2] != "Z"
    except (ProcessLookupError, FileNotFoundError):
        alive = False
    assert not alive                                      # the grandchild died with the group
    lib.mgl_destroy(pool)


def test_worker_pool_device_assignment(monkeypatch, tmp_path):
    """CUDA_VISIBLE_DEVICES / MAGGY_LOCAL_RANK per slot: 1 GPU/worker, gangs, and shared-device jobs."""
    seen = []

    class FakeLib:
        def __getattr__(self, name):
            def f(*a):
                if name == "mgl_spawn":
                    env = dict(kv.split("=", 1) for kv in a[3].decode().split("\0") if "=" in kv)
                    seen.append((a[1], a[4].decode(), env.get("MAGGY_LOCAL_RANK")))
                    return 1234
                return 1
            return f

    monkeypatch.setattr(workerpool, "_launcher", lambda: FakeLib())
    for kw, expect in [
        (dict(num_workers=2, gpus_per_worker=1, use_gpus=True), [(0, "0", "0"), (1, "1", "0")]),
        (dict(num_workers=2, gpus_per_worker=2, use_gpus=True), [(0, "0,1", "0"), (1, "0,1", "1"), (2, "2,3", "0"), (3, "2,3", "1")]),
        (dict(num_workers=2, gpus_per_worker=1, use_gpus=True, share_devices=True), [(0, "0,1", "0"), (1, "0,1", "1")]),
        (dict(num_workers=2, gpus_per_worker=1, use_gpus=False), [(0, "-1", "0"), (1, "-1", "0")]),
    ]:
        seen.clear()
        workerpool.WorkerPool(segment="/x", spec_path="/x", log_dir=str(tmp_path), **kw).start()
        assert seen == expect, (kw, seen)
