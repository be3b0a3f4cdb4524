"""Searchspace / Trial behaviour (golden values from the reference test-suite: maggy/tests/test_trial.py:24-48,
test_searchspace.py:24-77) plus a differential test against the importable reference modules."""
import importlib.util
import json
import os
import random
import sys

import pytest

from maggy_b200 import Searchspace, Trial

REF = "/root/reference/maggy"


def _load_ref(name, rel):
    path = os.path.join(REF, rel)
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_searchspace_init_and_access():
    sp = Searchspace(argument_param=("DOUBLE", [1, 5]), param2=("integer", [3, 4]))
    assert sp.argument_param == [1, 5]
    assert sp.names() == {"argument_param": "DOUBLE", "param2": "INTEGER"}
    assert sp.keys() == ["argument_param", "param2"]
    assert sp.get("param2") == [3, 4] and sp.get("nope", 7) == 7
    assert "param2" in sp and "zzz" not in sp
    assert sp.to_dict() == {"argument_param": ("DOUBLE", [1, 5]), "param2": ("INTEGER", [3, 4])}
    assert json.loads(sp.json()) == {"argument_param": ["DOUBLE", [1, 5]], "param2": ["INTEGER", [3, 4]]}
    assert [i["name"] for i in sp] == ["argument_param", "param2"]
    assert [i["name"] for i in sp.items()] == ["argument_param", "param2"]


@pytest.mark.parametrize("name,spec,exc", [
    ("p", ("DOUBLE", [1]), AssertionError),
    ("p", ("DOUBLE", [5, 1]), AssertionError),
    ("p", ("DOUBLE", ["a", 2]), ValueError),
    ("p", ("INTEGER", [1.5, 2]), ValueError),
    ("p", ("CATEGORICAL", []), ValueError),
    ("p", ("WHATEVER", [1, 2]), ValueError),
    ("p", ("DOUBLE",), ValueError),
    ("p", 5, ValueError),
    ("add", ("DOUBLE", [1, 2]), ValueError),  # reserved attribute name
])
def test_searchspace_validation(name, spec, exc):
    sp = Searchspace()
    with pytest.raises(exc):
        sp.add(name, spec)


def test_searchspace_duplicate_name():
    sp = Searchspace(a=("DOUBLE", [0, 1]))
    with pytest.raises(ValueError):
        sp.add("a", ("DOUBLE", [0, 1]))


def test_random_sampling_in_bounds():
    sp = Searchspace(x=("DOUBLE", [-1.0, 1.0]), n=("INTEGER", [2, 4]), d=("DISCRETE", [1, 10, 100]),
                     c=("CATEGORICAL", ["a", "b"]))
    for p in sp.get_random_parameter_values(200):
        assert -1.0 <= p["x"] <= 1.0 and p["n"] in (2, 3, 4) and p["d"] in (1, 10, 100) and p["c"] in ("a", "b")
    with pytest.raises(ValueError):
        sp.get_random_parameter_values("3")


def test_transform_roundtrip_and_reference_equivalence():
    ref = _load_ref("ref_searchspace", "searchspace.py")
    spec = dict(x=("DOUBLE", [-3.0, 3.0]), n=("INTEGER", [2, 9]), c=("CATEGORICAL", ["r", "g", "b"]))
    mine, theirs = Searchspace(**spec), ref.Searchspace(**spec)
    rng = random.Random(0)
    for _ in range(50):
        row = [rng.uniform(-3, 3), rng.randint(2, 9), rng.choice(["r", "g", "b"])]
        for norm in (False, True):
            a, b = mine.transform(row, normalize_categorical=norm), theirs.transform(row, normalize_categorical=norm)
            assert [float(x) for x in a] == pytest.approx([float(x) for x in b])
            back = mine.inverse_transform(a, normalize_categorical=norm)
            assert back[1] == row[1] and back[2] == row[2] and back[0] == pytest.approx(row[0])
            assert back == theirs.inverse_transform(b, normalize_categorical=norm) or back[0] == pytest.approx(
                theirs.inverse_transform(b, normalize_categorical=norm)[0])
    assert mine.list_to_dict([1.0, 2, "r"]) == {"x": 1.0, "n": 2, "c": "r"}
    with pytest.raises(ValueError):
        mine.list_to_dict([1.0])
    assert Searchspace.dict_to_list({"x": 1, "y": 2}) == [1, 2]


def test_trial_golden_id_and_json_roundtrip():
    t = Trial({"param1": 5, "param2": "ada"})
    assert t.trial_id == "3d1cc9fdb1d4d001"   # reference maggy/tests/test_trial.py:30
    assert t.id64 == int("3d1cc9fdb1d4d001", 16) and Trial.id_from_int(t.id64) == t.trial_id
    assert t.status == Trial.PENDING and t.early_stop is False
    t.append_metric({"value": 0.5, "step": 0})
    assert t.append_metric({"value": 0.6, "step": 0}) is None      # duplicate step ignored
    assert t.append_metric({"value": None, "step": 1}) is None
    assert t.append_metric({"value": 0.7, "step": 1}) == 1
    t.final_metric, t.status, t.duration = 0.7, Trial.FINALIZED, 12
    d = json.loads(t.to_json())
    assert set(d) == {"__class__", "trial_type", "trial_id", "params", "status", "early_stop", "final_metric",
                      "metric_history", "step_history", "metric_dict", "duration", "info_dict"}
    t2 = Trial.from_json(t.to_json())
    assert (t2.trial_id, t2.params, t2.status, t2.metric_history, t2.metric_dict) == (
        t.trial_id, t.params, t.status, [0.5, 0.7], {0: 0.5, 1: 0.7})
    with pytest.raises(ValueError):
        Trial.from_json(json.dumps({"__class__": "Nope"}))
    with pytest.raises(ValueError):
        Trial({1: 2})
    with pytest.raises(ValueError):
        Trial._generate_id([1, 2])


def test_ablation_trial_id_ignores_callables():
    a = Trial({"ablated_feature": "f", "ablated_layer": "None", "model_function": lambda: 1}, trial_type="ablation")
    b = Trial({"ablated_feature": "f", "ablated_layer": "None", "model_function": lambda: 2}, trial_type="ablation")
    assert a.trial_id == b.trial_id
