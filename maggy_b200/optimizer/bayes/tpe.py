"""Tree-structured Parzen estimator, BOHB flavour (reference: maggy/optimizer/bayes/tpe.py:31-266).

Finished trials are split into the best ``gamma`` fraction ("good") and the rest ("bad"), each with at
least ``n_hparams + 1`` members; a product KDE is fitted to each set; ``n_samples`` candidates are drawn
around good observations (truncated normals widened by ``bw_factor`` for continuous dimensions,
resample-with-probability-bandwidth for categorical ones) and the candidate with the largest density
ratio ``l(x)/g(x)`` wins.
"""
from __future__ import annotations

import numpy as np
import scipy.stats as sps

from .base import BaseAsyncBO
from .surrogate import ProductKDE


class TPE(BaseAsyncBO):
    def __init__(self, gamma=0.15, n_samples=24, bw_estimation="normal_reference", bw_factor=3, **kwargs):
        super().__init__(**kwargs)
        if self.interim_results:
            raise ValueError("Using interim results to update surrogate model is only supported for GP for now, "
                             "got TPE. Set `interim_results`=False when intitializing the optimizer or use GP")
        self.gamma = gamma
        self.n_samples = n_samples
        self.bw_estimation = bw_estimation
        self.min_bw = 1e-3
        self.bw_factor = bw_factor

    def init_model(self):
        pass

    @staticmethod
    def _get_vartype(maggy_vartype):
        if maggy_vartype in ("DOUBLE", "INTEGER"):
            return "c"
        if maggy_vartype == "CATEGORICAL":
            return "u"
        raise NotImplementedError("Only cont vartypes are implemented yer")

    def _var_types(self):
        return "".join(self._get_vartype(hp["type"]) for hp in self.searchspace.items())

    def _split_trials(self, budget=0):
        trials = [t for t in self.final_store if self._budget_match(t, budget)]
        sign = -1.0 if self.direction == "max" else 1.0
        order = np.argsort([sign * t.final_metric for t in trials])
        d = len(self.searchspace.keys())
        n = len(trials)
        n_good = max(d + 1, int(self.gamma * n))
        n_bad = max(d + 1, int((1 - self.gamma) * n))
        rows = self._param_rows(trials)
        good = [rows[i] for i in order[:n_good]]
        bad = [rows[i] for i in order[n_good:n_good + n_bad]]
        return good, bad

    def update_model(self, budget=0):
        good, bad = self._split_trials(budget)
        d = len(self.searchspace.keys())
        if d >= len(good) or d >= len(bad):
            self._log(f"Not enough observations to build model with budget {budget} yet. n_good_hparams: "
                      f"{len(good)}, n_bad_hparams: {len(bad)}, n_hparmas: {d}")
            return
        self._log(f"Update Model with budget {budget}. n_good_hparams: {len(good)}, n_bad_hparams: {len(bad)}")
        vt = self._var_types()
        ncat = [len(hp["values"]) if hp["type"] == "CATEGORICAL" else 0 for hp in self.searchspace.items()]
        self.models[budget] = {
            "good": ProductKDE(self._encode(good), vt, n_categories=ncat, bw=self.bw_estimation),
            "bad": ProductKDE(self._encode(bad), vt, n_categories=ncat, bw=self.bw_estimation),
        }

    def sampling_routine(self, budget=0):
        kde_good, kde_bad = self.models[budget]["good"], self.models[budget]["bad"]
        specs = list(self.searchspace.items())
        best_ratio, best = -np.inf, None
        for _ in range(self.n_samples):
            centre = kde_good.data[np.random.randint(0, len(kde_good.data))]
            cand = []
            for mean, bw, spec in zip(centre, kde_good.bw, specs):
                if spec["type"] in ("DOUBLE", "INTEGER"):
                    bw = max(bw, self.min_bw) * self.bw_factor
                    cand.append(sps.truncnorm.rvs(-mean / bw, (1 - mean) / bw, loc=mean, scale=bw))
                elif np.random.rand() < (1 - bw):
                    cand.append(int(mean))
                else:
                    cand.append(np.random.randint(len(spec["values"])))
            ratio = self._calculate_ei(cand, kde_good, kde_bad)
            if ratio > best_ratio:
                best_ratio, best = ratio, cand
        return self.searchspace.list_to_dict(self.searchspace.inverse_transform(best))

    @staticmethod
    def _calculate_ei(x, kde_good, kde_bad):
        return max(1e-32, kde_good.pdf(x)) / max(kde_bad.pdf(x), 1e-32)
