"""Run the UNMODIFIED reference's ``lagom(train_fn, BaseConfig)`` in a scratch directory and print its result as JSON
(helper of tests/test_reference_differential.py; needs baseline/_ref)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baseline import refenv  # noqa: E402

refenv.install()
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
os.chdir(sys.argv[1])
mode = sys.argv[2]
from maggy.config import BaseConfig  # noqa: E402
from maggy.experiment import experiment  # noqa: E402


def train_dict():
    return {"accuracy": 0.9, "loss": 0.1}


def train_scalar():
    return 0.25


res = experiment.lagom(train_dict if mode == "dict" else train_scalar, BaseConfig(name="diff", description="d"))
print("RESULT " + json.dumps(res, default=str))
