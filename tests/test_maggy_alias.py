"""The ``maggy`` alias package: reference-style imports resolve to the maggy_b200 modules (same objects)."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import sys
sys.path.insert(0, {root!r})
from maggy import experiment, Searchspace          # reference: maggy/__init__.py:17-21
from maggy.optimizer import Asha, RandomSearch
from maggy.optimizer.bayes import GP, TPE
from maggy.experiment_config import OptimizationConfig   # legacy notebooks
from maggy.config import HyperparameterOptConfig, AblationConfig, TorchDistributedConfig, TfDistributedConfig
from maggy.ablation import AblationStudy
from maggy.earlystop import AbstractEarlyStop, MedianStoppingRule
from maggy.core.exceptions import EarlyStopException
from maggy.callbacks import KerasBatchEnd, KerasEpochEnd
import maggy.tensorboard as tb
import maggy, maggy_b200, maggy_b200.earlystop
assert Searchspace is maggy_b200.Searchspace
assert sys.modules["maggy.earlystop"] is maggy_b200.earlystop
assert OptimizationConfig is HyperparameterOptConfig
assert callable(experiment.lagom) and callable(tb.logdir)
try:
    import maggy.does_not_exist
except ImportError:
    pass
else:
    raise SystemExit("bogus submodule imported")
print("ok")
"""


def test_reference_style_imports_resolve_to_maggy_b200(tmp_path):
    res = subprocess.run([sys.executable, "-c", CODE.format(root=ROOT)], capture_output=True, text=True, cwd=tmp_path,
                         timeout=120)
    assert res.returncode == 0, res.stderr
    assert res.stdout.strip().endswith("ok")
