"""Search policies: pruner schedules (golden tables), ASHA, grid/random, Bayesian optimizers, median rule
(Python oracle == reference module == native policy engine)."""
import importlib.util
import os
import random
import tempfile

import numpy as np
import pytest

from maggy_b200 import Searchspace, Trial
from maggy_b200.earlystop import MedianStoppingRule, NoStoppingRule
from maggy_b200.optimizer import Asha, GridSearch, RandomSearch, SingleRun
from maggy_b200.optimizer.bayes import GP, TPE
from maggy_b200.optimizer.bayes.surrogate import GaussianProcess, ProductKDE
from maggy_b200.pruner import Hyperband


def _wire(opt, sp, n, direction="max"):
    opt.searchspace, opt.num_trials, opt.trial_store, opt.final_store, opt.direction = sp, n, {}, [], direction
    opt._initialize(tempfile.mkdtemp())
    return opt


def _finish(opt, t, metric):
    t.final_metric, t.metric_history, t.status = metric, [metric], Trial.FINALIZED
    opt.trial_store.pop(t.trial_id, None)
    opt.final_store.append(t)


@pytest.mark.parametrize("args,budgets,brackets,total", [
    ((1, 9, 3, 3), [1, 3, 9], [[9, 3, 1], [3, 1], [3]], 20),
    ((1, 27, 3, 4), [1, 3, 9, 27], [[27, 9, 3, 1], [9, 3, 1], [6, 2], [4]], 65),
])
def test_hyperband_golden_schedule(args, budgets, brackets, total):
    hb = Hyperband(*args, trial_metric_getter=lambda ids: {})
    assert hb.budgets == budgets
    assert [b.n_configs for b in hb.iterations] == brackets
    assert hb.num_trials() == total


def test_hyperband_matches_reference_run():
    path = "/root/reference/maggy/pruner/hyperband.py"
    if not os.path.exists(path):
        pytest.skip("reference not mounted")
    # the reference module imports maggy.pruner.abstractpruner -> maggy.core.environment (TF-free): load by path
    import sys
    import types

    pkg = types.ModuleType("maggy"); pkg.__path__ = ["/root/reference/maggy"]
    saved = {k: sys.modules.get(k) for k in ("maggy",)}
    try:
        sys.modules["maggy"] = pkg
        try:
            from maggy.pruner.hyperband import Hyperband as RefHB  # noqa
        except Exception as e:  # pragma: no cover
            pytest.skip(f"reference pruner not importable here: {e}")
        scores = {}

        def run(cls):
            rng = random.Random(5)
            done = {}
            hb = cls(1, 9, 3, 3, trial_metric_getter=lambda ids: (
                {i: done[i] for i in ([ids] if isinstance(ids, str) else ids) if i in done}))
            if cls is Hyperband:
                hb.initialize_logger(tempfile.mkdtemp())
            else:  # the reference logger needs its TF-importing environment: give it an already-closed file
                hb.fd = types.SimpleNamespace(closed=True)
            order, n = [], 0
            while True:
                r = hb.pruning_routine()
                if r is None:
                    break
                if r == "IDLE":
                    continue
                n += 1
                tid = f"t{n}"
                hb.report_trial(r["trial_id"], tid)
                done[tid] = rng.random()
                order.append((r["trial_id"] is None, r["budget"]))
            return order

        assert run(Hyperband) == run(RefHB)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith("maggy.")]:
            sys.modules.pop(k, None)


def test_hyperband_full_run_promotes_best():
    done = {}

    def getter(ids):
        ids = [ids] if isinstance(ids, str) else ids
        return {i: done[i] for i in ids if i in done}

    hb = Hyperband(1, 9, 3, 1, trial_metric_getter=getter)
    runs, n = [], 0
    while True:
        r = hb.pruning_routine()
        if r is None:
            break
        assert r != "IDLE"
        n += 1
        tid = f"t{n:02d}"
        hb.report_trial(r["trial_id"], tid)
        parent_score = done.get(r["trial_id"], n)  # promoted runs inherit their parent's quality
        done[tid] = parent_score
        runs.append((tid, r["trial_id"], r["budget"]))
    assert [b for _, _, b in runs] == [1] * 9 + [3] * 3 + [9]
    assert {p for _, p, b in runs if b == 3} == {"t01", "t02", "t03"}   # lowest score = best
    assert hb.finished()


def test_random_and_grid_search():
    sp = Searchspace(x=("DOUBLE", [0.0, 1.0]), c=("CATEGORICAL", ["a", "b"]))
    rs = _wire(RandomSearch(), sp, 5)
    seen = []
    while (t := rs.get_suggestion()) is not None:
        seen.append(t)
    assert len(seen) == 5 and len({t.trial_id for t in seen}) == 5
    assert all(t.info_dict["sample_type"] == "random" for t in seen)
    with pytest.raises(NotImplementedError):
        _wire(RandomSearch(), Searchspace(c=("CATEGORICAL", ["a", "b"])), 2)
    gsp = Searchspace(a=("DISCRETE", [1, 2, 3]), b=("CATEGORICAL", ["x", "y"]))
    assert GridSearch.get_num_trials(gsp) == 6
    gs = _wire(GridSearch(), gsp, 6)
    combos = set()
    while (t := gs.get_suggestion()) is not None:
        combos.add((t.params["a"], t.params["b"]))
    assert combos == {(a, b) for a in (1, 2, 3) for b in ("x", "y")}
    with pytest.raises(NotImplementedError):
        GridSearch.get_num_trials(sp)
    sr = _wire(SingleRun(), Searchspace(), 3)
    assert len({sr.get_suggestion().trial_id for _ in range(3)}) == 3 and sr.get_suggestion() is None


def test_random_search_with_hyperband_pruner():
    sp = Searchspace(x=("DOUBLE", [0.0, 1.0]))
    rs = RandomSearch(pruner="hyperband", pruner_kwargs=dict(min_budget=1, max_budget=9, eta=3, n_iterations=1))
    _wire(rs, sp, rs.pruner.num_trials(), "max")
    budgets = []
    last = None
    while True:
        t = rs.get_suggestion(last)
        if t is None:
            break
        assert t != "IDLE"
        budgets.append(t.params["budget"])
        rs.trial_store[t.trial_id] = t
        _finish(rs, t, t.params["x"])
        last = t
    assert budgets == [1] * 9 + [3] * 3 + [9]
    top = max(rs.final_store, key=lambda t: (t.params["budget"]))
    best_x = max(t.params["x"] for t in rs.final_store if t.params["budget"] == 1)
    assert top.params["x"] == pytest.approx(best_x)  # the best rung-0 config reached the top rung


@pytest.mark.parametrize("direction", ["max", "min"])
def test_asha_promotions_respect_direction(direction):
    sp = Searchspace(x=("DOUBLE", [0.0, 1.0]))
    asha = _wire(Asha(reduction_factor=2, resource_min=1, resource_max=4), sp, 8, direction)
    assert asha.max_rung == 2
    first = asha.get_suggestion(None)
    assert first.params["budget"] == 1
    running, n_promoted = [first], 0
    for step in range(60):
        t = running.pop(0)
        _finish(asha, t, t.params["x"])
        nxt = asha.get_suggestion(t)
        if nxt is None:
            break
        if nxt.params["budget"] > 1:
            n_promoted += 1
            rung = {2: 0, 4: 1}[nxt.params["budget"]]
            finished = sorted((u.params["x"] for u in asha.rungs[rung] if u.status == Trial.FINALIZED),
                              reverse=(direction == "max"))
            top = finished[: len(finished) // 2]
            assert nxt.params["x"] in top  # promoted config was in the best 1/rf of its rung at that time
        running.append(nxt)
    assert n_promoted > 0, "nothing was promoted"
    assert 2 in asha.rungs  # ended because a config reached the top rung
    with pytest.raises(Exception):
        Asha(reduction_factor=1)
    with pytest.raises(Exception):
        Asha(resource_min=4, resource_max=2)


def _optimize(opt, n, direction, seed=0):
    np.random.seed(seed)
    random.seed(seed)
    sp = Searchspace(x=("DOUBLE", [-3.0, 3.0]), y=("DOUBLE", [-3.0, 3.0]), c=("CATEGORICAL", ["a", "b"]))
    _wire(opt, sp, n, direction)
    f = lambda p: (p["x"] - 1) ** 2 + (p["y"] + 0.5) ** 2 + (0.5 if p["c"] == "b" else 0)  # noqa: E731
    last = None
    while True:
        t = opt.get_suggestion(last)
        if t is None:
            break
        opt.trial_store[t.trial_id] = t
        _finish(opt, t, f(t.params) if direction == "min" else -f(t.params))
        last = t
    return min(abs(t.final_metric) for t in opt.final_store), opt


def test_gp_and_tpe_beat_random_search():
    rand_best, _ = _optimize(RandomSearch(), 40, "min")
    gp_best, gp = _optimize(GP(num_warmup_trials=8, acq_optimizer_kwargs={"n_points": 1500}), 40, "min")
    assert gp_best < rand_best and gp_best < 0.05
    assert len(gp.final_store) == 40 and len({t.trial_id for t in gp.final_store}) == 40
    # model-guided samples must be (much) better on average than the random ones of the same runs
    for cls, kw, direction in ((TPE, dict(num_warmup_trials=10), "max"),
                               (GP, dict(num_warmup_trials=8, acq_optimizer_kwargs={"n_points": 800}), "min")):
        model, rnd = [], []
        for seed in range(3):
            _, opt = _optimize(cls(**kw), 50, direction, seed=seed)
            model += [abs(t.final_metric) for t in opt.final_store if t.info_dict["sample_type"] == "model"]
            rnd += [abs(t.final_metric) for t in opt.final_store if t.info_dict["sample_type"] != "model"]
        assert len(model) > 20 and np.mean(model) < 0.5 * np.mean(rnd), (cls.__name__, np.mean(model), np.mean(rnd))


def test_gp_options_and_validation():
    for kw in (dict(async_strategy="asy_ts"), dict(acq_fun="LCB"), dict(acq_fun="PI", acq_optimizer="sampling"),
               dict(impute_strategy="kb"), dict(impute_strategy="cl_mean")):
        best, _ = _optimize(GP(num_warmup_trials=6, acq_optimizer_kwargs={"n_points": 300}, **kw), 16, "min")
        assert best < 3.0
    with pytest.raises(ValueError):
        GP(async_strategy="nope")
    with pytest.raises(ValueError):
        GP(acq_fun="AsyTS")
    with pytest.raises(ValueError):
        GP(impute_strategy="zzz")
    with pytest.raises(ValueError):
        GP(acq_optimizer="adam")
    with pytest.raises(ValueError):
        TPE(interim_results=True)
    with pytest.raises(ValueError):
        _wire(GP(), Searchspace(d=("DISCRETE", [1, 2])), 4)


def test_gp_surrogate_fits_and_gradients():
    rng = np.random.RandomState(0)
    X = rng.rand(30, 2)
    y = np.sin(3 * X[:, 0]) + X[:, 1] ** 2
    gp = GaussianProcess(2, random_state=0).fit(X, y)
    mu, sd = gp.predict(X, return_std=True)
    assert np.abs(mu - y).max() < 0.05 and (sd >= 0).all()
    x0 = np.array([0.3, 0.6])
    m, s, dm, ds = gp.predict(x0.reshape(1, -1), return_std=True, return_mean_grad=True, return_std_grad=True)
    eps = 1e-5
    for j in range(2):
        e = np.zeros(2); e[j] = eps
        mp, sp = gp.predict((x0 + e).reshape(1, -1), return_std=True)
        mm, sm = gp.predict((x0 - e).reshape(1, -1), return_std=True)
        assert dm[j] == pytest.approx((mp[0] - mm[0]) / (2 * eps), rel=1e-3, abs=1e-5)
        assert ds[j] == pytest.approx((sp[0] - sm[0]) / (2 * eps), rel=1e-2, abs=1e-5)
    assert gp.sample_y(X[:5], n_samples=3).shape == (5, 3)


def test_product_kde():
    rng = np.random.RandomState(1)
    data = np.c_[rng.normal(0.5, 0.1, 200), rng.randint(0, 3, 200)]
    kde = ProductKDE(data, "cu", n_categories=[0, 3])
    assert kde.pdf([0.5, 1]) > kde.pdf([0.9, 1]) > 0
    # continuous marginal integrates to ~1 for a fixed category probability mass
    xs = np.linspace(-0.5, 1.5, 400)
    total = sum(np.trapz([kde.pdf([x, c]) for x in xs], xs) for c in range(3))
    assert total == pytest.approx(1.0, abs=0.02)


def _mk_trial(hist):
    t = Trial({"h": str(hist)})
    t.metric_history = list(hist)
    return t


def test_median_rule_matches_reference():
    path = "/root/reference/maggy/earlystop/medianrule.py"
    rng = random.Random(3)
    finals = [_mk_trial([rng.random() for _ in range(rng.randint(1, 8))]) for _ in range(12)]
    ref_check = None
    if os.path.exists(path):
        import sys
        import types

        pkg = types.ModuleType("maggy"); pkg.__path__ = ["/root/reference/maggy"]
        sys.modules["maggy"] = pkg
        try:
            from maggy.earlystop.medianrule import MedianStoppingRule as RefRule

            ref_check = RefRule.earlystop_check
        except Exception:
            ref_check = None
        finally:
            for k in [k for k in sys.modules if k == "maggy" or k.startswith("maggy.")]:
                sys.modules.pop(k, None)
    for _ in range(60):
        cand = _mk_trial([rng.random() for _ in range(rng.randint(1, 8))])
        for direction in ("max", "min"):
            mine = MedianStoppingRule.earlystop_check(cand, finals, direction)
            if ref_check is not None:
                try:
                    theirs = ref_check(cand, finals, direction)
                except Exception:
                    theirs = None  # the reference raises when no finished trial is long enough
                assert mine == theirs
    assert NoStoppingRule.earlystop_check(finals[0], finals, "max") is None
    assert MedianStoppingRule.earlystop_check(_mk_trial([]), finals, "max") is None


def test_loco_feature_store_dataset_generator_drops_the_ablated_feature(tmp_path, monkeypatch):
    """Without a custom generator LOCO reads the study's training dataset (parquet here) and leaves the ablated
    feature's column out (reference loco.py:41-80 does it through hsfs / TFRecords)."""
    import numpy as np
    import pandas as pd

    from maggy_b200.ablation import AblationStudy
    from maggy_b200.ablation.ablator.loco import LOCO
    from maggy_b200.core.environment.singleton import EnvSing

    root = tmp_path / "Training_Datasets" / "titanic_2"
    root.mkdir(parents=True)
    rng = np.random.default_rng(0)
    df = pd.DataFrame({"age": rng.random(10), "fare": rng.random(10), "sex": rng.integers(0, 2, 10).astype(float),
                       "survived": rng.integers(0, 2, 10)})
    df.iloc[:6].to_parquet(root / "part-0.parquet")
    df.iloc[6:].to_parquet(root / "part-1.parquet")
    monkeypatch.setenv("MAGGY_TRAINING_DATASETS", str(tmp_path / "Training_Datasets"))
    monkeypatch.setenv("MAGGY_LOG_DIR", str(tmp_path / "log"))
    EnvSing.reset()
    study = AblationStudy("titanic", 2, "survived")
    study.features.include("fare")
    study.model.set_base_model_generator(lambda: None)
    loco = LOCO(study, None)
    full = loco.get_dataset_generator()(num_epochs=1, batch_size=4, shuffle=False)
    assert full.feature_names == ["age", "fare", "sex"] and len(full) == 3
    ablated = loco.get_dataset_generator(ablated_feature="fare")(num_epochs=2, batch_size=4, shuffle=False)
    batches = list(ablated)
    assert ablated.feature_names == ["age", "sex"] and len(batches) == 6
    x0, y0 = batches[0]
    assert tuple(x0.shape) == (4, 2) and tuple(y0.shape) == (4,) and str(y0.dtype) == "torch.int64"
    np.testing.assert_allclose(x0.numpy(), df[["age", "sex"]].to_numpy(dtype=np.float32)[:4])
    with pytest.raises(KeyError):
        loco.get_dataset_generator(ablated_feature="nope")(1, 4)
    EnvSing.reset()


def test_device_asha_tables_agree_with_the_python_optimizer():
    """csrc/include/mg_asha.h (the rung tables K9 keeps on the GPU; exercised here through the host library) takes the
    same promote / fresh-configuration decisions as maggy_b200.optimizer.Asha on a random stream of results."""
    import ctypes
    import random

    from maggy_b200 import Searchspace
    from maggy_b200.core import rpc
    from maggy_b200.optimizer import Asha
    from maggy_b200.trial import Trial

    lib = rpc._host()
    for direction, rf, rmax in (("max", 2, 4), ("min", 3, 27), ("max", 2, 16)):
        opt = Asha(reduction_factor=rf, resource_min=1, resource_max=rmax)
        opt.searchspace = Searchspace(x=("DOUBLE", [0.0, 1.0]))
        opt.searchspace.seed(1)
        opt.num_trials, opt.direction = 4096, direction
        opt.trial_store, opt.final_store = {}, []
        opt.initialize()
        opt.max_rung = 6   # far-away closing rung: compare a long decision stream
        tab = lib.mgr_asha_create(rf, opt.max_rung, 512, 1 if direction == "max" else -1)
        rng = random.Random(7)
        running = [opt.get_suggestion() for _ in range(5)]
        promotions = 0
        for step in range(300):
            t = running.pop(rng.randrange(len(running)))
            t.status, t.final_metric = Trial.FINALIZED, rng.random()
            parent, rung = ctypes.c_uint64(), ctypes.c_uint32()
            kind = lib.mgr_asha_final(tab, opt.rung_of(t), int(t.trial_id, 16), t.final_metric, ctypes.byref(parent),
                                      ctypes.byref(rung))
            nxt = opt.get_suggestion(t)
            if nxt is None:   # a configuration reached the closing rung
                break
            if nxt.info_dict["sample_type"] == "promoted":
                promotions += 1
                assert kind == Asha.PROMOTE, (direction, step)
                assert f"{parent.value:016x}" == nxt.info_dict["parent"] and opt.rung_of(nxt) == rung.value + 1
            else:
                assert kind == Asha.NEW, (direction, step)
            running.append(nxt)
        assert promotions > 20 and step > 100
        lib.mgr_asha_free(tab)


def test_asha_applies_device_decisions_and_falls_back():
    from maggy_b200 import Searchspace
    from maggy_b200.optimizer import Asha
    from maggy_b200.trial import Trial

    opt = Asha(reduction_factor=2, resource_min=1, resource_max=4)
    opt.searchspace = Searchspace(x=("DOUBLE", [0.0, 1.0]))
    opt.num_trials, opt.direction = 64, "max"
    opt.trial_store, opt.final_store = {}, []
    opt.initialize()
    a, b = opt.get_suggestion(), opt.get_suggestion()
    for t, m in ((a, 0.3), (b, 0.9)):
        t.status, t.final_metric = Trial.FINALIZED, m
    child = opt.apply_device_decision(Asha.PROMOTE, 0, b.trial_id, b)
    assert child.params["budget"] == 2 and child.info_dict["parent"] == b.trial_id and b.trial_id in opt.promoted[0]
    assert opt.rung_of(child) == 1 and child in opt.rungs[1]
    again = opt.apply_device_decision(Asha.PROMOTE, 0, b.trial_id, b)          # already promoted: host decides instead
    assert again.info_dict["sample_type"] == "random"
    unknown = opt.apply_device_decision(Asha.PROMOTE, 0, "00000000deadbeef", b)   # parent the host never placed
    assert unknown.info_dict["sample_type"] == "random"
    fresh = opt.apply_device_decision(Asha.NEW, 0, "0" * 16, a)
    assert fresh.params["budget"] == 1 and fresh in opt.rungs[0]
