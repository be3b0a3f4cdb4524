"""Gaussian-process Bayesian optimization (reference: maggy/optimizer/bayes/gp.py:34-373).

Surrogate: ARD Matern-5/2 GP with learned amplitude and noise, normalised targets
(:class:`~maggy_b200.optimizer.bayes.surrogate.GaussianProcess`). Asynchrony: ``impute`` (constant liar
``cl_min/cl_max/cl_mean`` or kriging believer ``kb`` for configurations still running) with EI / LCB / PI,
or ``asy_ts`` (asynchronous Thompson sampling). The acquisition is minimised over ``n_points`` random
candidates and, with ``acq_optimizer="lbfgs"``, polished from the best ``n_restarts_optimizer`` of them.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import fmin_l_bfgs_b

from .acquisitions import AsyTS, GaussianProcess_EI, GaussianProcess_LCB, GaussianProcess_PI
from .base import BaseAsyncBO
from .surrogate import GaussianProcess

_ACQ = {"impute": {"EI": GaussianProcess_EI, "LCB": GaussianProcess_LCB, "PI": GaussianProcess_PI},
        "asy_ts": {"AsyTS": AsyTS}}
_IMPUTE = ("cl_min", "cl_max", "cl_mean", "kb")


class GP(BaseAsyncBO):
    def __init__(self, async_strategy="impute", impute_strategy="cl_min", acq_fun=None, acq_fun_kwargs=None,
                 acq_optimizer="lbfgs", acq_optimizer_kwargs=None, **kwargs):
        super().__init__(**kwargs)
        if async_strategy not in _ACQ:
            raise ValueError(f"Expected async_strategy to be in {list(_ACQ)} with GP as surrogate, got {async_strategy}")
        if async_strategy == "impute" and self.pruner and not self.interim_results:
            raise ValueError("Optimizer GP with async strategy `impute` only supports Pruner with "
                             f"interim_results==True, got {self.interim_results}")
        choices = _ACQ[async_strategy]
        if acq_fun is not None and acq_fun not in choices:
            raise ValueError(f"Expected acq_fun to be in {list(choices)} with GP as surrogate and {async_strategy} "
                             f"as async_strategy, got {acq_fun}")
        self.async_strategy = async_strategy
        self.acq_fun = choices[acq_fun or next(iter(choices))]()
        self.acq_func_kwargs = acq_fun_kwargs
        if acq_optimizer not in ("sampling", "lbfgs"):
            raise ValueError(f"expected acq_optimizer to be in ['sampling', 'lbfgs'], got {acq_optimizer}")
        self.acq_optimizer = acq_optimizer
        opt_kw = acq_optimizer_kwargs or {}
        if async_strategy == "asy_ts":
            self.n_points = int(np.clip(opt_kw.get("n_points", 100), 10, 1000))
        else:
            self.n_points = int(opt_kw.get("n_points", 10000))
        self.n_restarts_optimizer = opt_kw.get("n_restarts_optimizer", 5)
        self.acq_optimizer_kwargs = opt_kw
        if async_strategy == "impute":
            if impute_strategy not in _IMPUTE:
                raise ValueError(f"expected impute_strategy to be in {list(_IMPUTE)}, got {impute_strategy}")
            self.impute_strategy = impute_strategy
        self.base_model = None

    # ------------------------------------------------------------------ surrogate
    def init_model(self):
        n_dims = len(self.searchspace.keys()) + (1 if self.interim_results else 0)
        self.base_model = GaussianProcess(n_dims, amp_bounds=(0.01, 1000.0), ls_bounds=(0.01, 100.0),
                                          n_restarts_optimizer=2, normalize_y=True)

    def update_model(self, budget=0):
        self._log(f"start updateing model with budget {budget}")
        n_obs = len([t for t in self.final_store if self._budget_match(t, budget)])
        if len(self.searchspace.keys()) > n_obs:
            self._log(f"not enough observations available to build with budget {budget} yet. At least "
                      f"{len(self.searchspace.keys())} needed, got {n_obs}")
            return
        X, y = self.get_XY(budget=budget, interim_results=self.interim_results,
                           interim_results_interval=self.interim_results_interval)
        model = self.base_model.clone()
        model.fit(X, y)
        self._log("fitted model with data")
        self.models[budget] = model

    def sampling_routine(self, budget=0):
        sp = self.searchspace
        cands = sp.get_random_parameter_values(self.n_points)
        X = np.array([sp.transform(sp.dict_to_list(c), normalize_categorical=True) for c in cands], dtype=float)
        if self.interim_results:
            X = np.append(X, np.ones((X.shape[0], 1)), 1)  # ask for the prediction at the full budget
        y_opt = self.ybest(budget)
        model = self.models[budget]
        values = self.acq_fun.evaluate(X=X, surrogate_model=model, y_opt=y_opt, acq_func_kwargs=self.acq_func_kwargs)
        if self.acq_optimizer == "sampling":
            best_x = X[int(np.argmin(values))]
        else:
            approx_grad = self.async_strategy == "asy_ts"
            bounds = [(0.0, 1.0)] * X.shape[1]
            best_x, best_val = None, np.inf
            for x0 in X[np.argsort(values)[: self.n_restarts_optimizer]]:
                x, val, _ = fmin_l_bfgs_b(func=self.acq_fun.evaluate_1_d, x0=x0,
                                          args=(model, y_opt, self.acq_func_kwargs), bounds=bounds,
                                          approx_grad=approx_grad, maxiter=20)
                if val < best_val:
                    best_x, best_val = x, val
        best_x = np.clip(best_x, 0.0, 1.0)[:len(sp.keys())]
        return sp.list_to_dict(sp.inverse_transform(best_x, normalize_categorical=True))

    # ------------------------------------------------------------------ async imputation
    def impute_metric(self, hparams, budget=0):
        """Metric to pretend for a configuration that is still running (returned in the user's sign
        convention, like a real final metric)."""
        if self.impute_strategy == "cl_min":
            val = self.ybest(budget)
        elif self.impute_strategy == "cl_max":
            val = self.yworst(budget)
        elif self.impute_strategy == "cl_mean":
            val = self.ymean(budget)
        elif self.impute_strategy == "kb":
            sp = self.searchspace
            x = sp.transform([hparams[k] for k in sp.keys()], normalize_categorical=True)
            if self.interim_results:
                x = np.append(x, 1)
            val = self.models[budget].predict(np.array(x, dtype=float).reshape(1, -1))[0]
        else:
            raise NotImplementedError(f"cl_strategy {self.impute_strategy} is not implemented, please choose from "
                                      "('cl_min', 'cl_max', 'cl_mean')")
        # ybest/yworst/ymean and the model live in "lower is better" space; get_XY applies the sign again
        return -val if self.direction == "max" else val
