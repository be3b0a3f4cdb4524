"""ThreadSanitizer build + stress run of the native control rings (csrc/ctrl/ring.cpp)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, sanitize):
    exe = str(tmp_path / ("ring_stress_tsan" if sanitize else "ring_stress"))
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-I", os.path.join(ROOT, "maggy_b200", "csrc", "include"),
           os.path.join(ROOT, "maggy_b200", "csrc", "ctrl", "ring.cpp"),
           os.path.join(ROOT, "tests", "native", "ring_stress.cpp"), "-o", exe, "-lpthread", "-lrt"]
    if sanitize:
        cmd.insert(1, "-fsanitize=thread")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    return exe, res


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_ring_stress_plain(tmp_path):
    exe, res = _build(tmp_path, sanitize=False)
    assert res.returncode == 0, res.stderr[-3000:]
    run = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "RING_STRESS OK" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_ring_stress_under_thread_sanitizer(tmp_path):
    exe, res = _build(tmp_path, sanitize=True)
    if res.returncode != 0 and "tsan" in (res.stderr or "").lower():
        pytest.skip("libtsan is not installed")
    assert res.returncode == 0, res.stderr[-3000:]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66 second_deadlock_stack=1")
    run = subprocess.run([exe, "3000"], capture_output=True, text=True, timeout=600, env=env)
    assert "ThreadSanitizer" not in run.stderr, run.stderr[-4000:]
    assert run.returncode == 0 and "RING_STRESS OK" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]
