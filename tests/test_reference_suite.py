"""The reference's OWN unit tests (maggy/tests/*.py), unmodified, run against this repository's ``maggy`` alias package.

Spark is replaced by a two-executor ``sc`` fixture stub and ``tensorflow`` by an empty stand-in module (only imported at
module scope by the tests that run). Not run: the two tests that call ``RandomSearch(5, sp, [])`` — a constructor the
reference itself no longer has (they raise the same TypeError against the reference) — and the two end-to-end tests that
need a live Spark cluster / MNIST download. Skipped when the reference tree is absent."""
import os
import shutil
import subprocess
import sys

import pytest

REF_TESTS = os.environ.get("MAGGY_REFERENCE_TESTS", "/root/reference/maggy/tests")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFTEST = '''
import sys, types
import pytest

tf = types.ModuleType("tensorflow")
tf.keras = types.ModuleType("tensorflow.keras")
sys.modules.setdefault("tensorflow", tf)
sys.modules.setdefault("tensorflow.keras", tf.keras)


class _Conf:
    def get(self, key, default=None):
        return {"spark.executor.instances": "2"}.get(key, default)


class _SC:
    _conf = _Conf()


@pytest.fixture
def sc():
    return _SC()
'''


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not available")
def test_reference_unit_tests_pass_against_the_alias_package(tmp_path):
    for f in ("test_trial.py", "test_searchspace.py", "test_maggy.py", "test_randomsearch.py"):
        shutil.copy(os.path.join(REF_TESTS, f), tmp_path / f)
    (tmp_path / "conftest.py").write_text(CONFTEST)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), MAGGY_NUM_EXECUTORS="2")
    deselect = []
    for t in ("test_randomsearch.py::test_randomsearch_init", "test_randomsearch.py::test_randomsearch_initialize",
              "test_randomsearch.py::test_randomsearch", "test_randomsearch.py::test_dt_tensorflow"):
        deselect += ["--deselect", t]
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "."] + deselect,
                         cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    tail = out.stdout[-1500:] + out.stderr[-500:]
    assert out.returncode == 0, tail
    assert "7 passed" in out.stdout, tail
