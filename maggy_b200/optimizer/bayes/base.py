"""Skeleton of asynchronous Bayesian optimization (reference: maggy/optimizer/bayes/base.py:26-680).

Flow of ``get_suggestion``: finished? -> pruner decides budget / promotion -> warm-up buffer ->
with probability ``random_fraction`` a random config -> otherwise (re)fit the surrogate for the
relevant budget and sample from it -> reject duplicates (up to 3 forced-random retries).
Async handling: configurations that are still being evaluated ("busy locations") can be imputed with a
constant liar / kriging believer value so the surrogate does not propose them again.
"""
from __future__ import annotations

import time
from abc import abstractmethod
from copy import deepcopy

import numpy as np

from ..abstractoptimizer import AbstractOptimizer


class BaseAsyncBO(AbstractOptimizer):
    def __init__(self, num_warmup_trials=15, random_fraction=0.33, interim_results=False,
                 interim_results_interval=10, **kwargs):
        super().__init__(**kwargs)
        self.num_warmup_trials = num_warmup_trials
        self.warmup_sampling = "random"
        self.warmup_configs = []
        self.models = {}  # budget -> fitted surrogate
        self.random_fraction = random_fraction
        self.interim_results = interim_results
        self.interim_results_interval = interim_results_interval
        self.sampling_time_start = 0.0
        # TPE's categorical kernel works on raw category indices, the GP wants everything in [0, 1]
        self.normalize_categorical = self.name() != "TPE"

    # ------------------------------------------------------------------ surrogate hooks
    @abstractmethod
    def init_model(self):
        ...

    @abstractmethod
    def update_model(self, budget=0):
        ...

    @abstractmethod
    def sampling_routine(self, budget=0):
        ...

    # ------------------------------------------------------------------ AbstractOptimizer
    def initialize(self):
        sp = self.searchspace
        continuous = False
        for hp in sp.items():
            if hp["type"] == sp.DISCRETE:
                raise ValueError("This version of Bayesian Optimization does not support DISCRETE Hyperparameters "
                                 f"yet, please encode {hp['name']} as INTEGER")
            continuous |= hp["type"] in (sp.DOUBLE, sp.INTEGER)
        if not continuous:
            raise ValueError("In this version of Bayesian Optimization at least one hparam has to be continuous "
                             "(DOUBLE or INTEGER)")
        self.warmup_routine()
        self.init_model()

    def warmup_routine(self):
        if self.warmup_sampling != "random":
            raise NotImplementedError(f"warmup sampling {self.warmup_sampling} doesnt exist, use random")
        self.warmup_configs = self.searchspace.get_random_parameter_values(self.num_warmup_trials)

    def _experiment_finished(self) -> bool:
        if self.pruner:
            return bool(self.pruner.finished())
        if len(self.final_store) >= self.num_trials:
            self._log(f"Finished experiment, ran {len(self.final_store)}/{self.num_trials} trials")
            return True
        return False

    def _random_trial(self, run_budget, sample_type="random"):
        hparams = self.searchspace.get_random_parameter_values(1)[0]
        return self.create_trial(hparams=hparams, sample_type=sample_type, run_budget=run_budget)

    def get_suggestion(self, trial=None):
        self._log("### start get_suggestion ###")
        self.sampling_time_start = time.time()
        self._log(f"last finished trial: {trial.trial_id} with params {trial.params}" if trial
                  else "no previous finished trial")
        if self._experiment_finished():
            return None

        run_budget = model_budget = 0
        if self.pruner:
            run = self.pruner.pruning_routine()
            if run == "IDLE":
                self._log("Worker is IDLE and has to wait until a new trial can be scheduled")
                return "IDLE"
            if run is None:
                self._log("Experiment has finished")
                return None
            if run["trial_id"]:
                parent = run["trial_id"]
                hparams = deepcopy(self.get_hparams_dict(trial_ids=parent)[parent])
                nxt = self.create_trial(hparams=hparams, sample_type="promoted", run_budget=run["budget"])
                self.pruner.report_trial(original_trial_id=parent, new_trial_id=nxt.trial_id)
                self._log(f"use hparams from promoted trial {parent}. \n start trial {nxt.trial_id}: {nxt.params} \n")
                return nxt
            run_budget = run["budget"]
            model_budget = 0 if self.interim_results else run_budget

        if self.warmup_configs:
            self._log("take sample from warmup buffer")
            nxt = self.create_trial(hparams=self.warmup_configs.pop(), sample_type="random", run_budget=run_budget)
        elif np.random.rand() < self.random_fraction:
            nxt = self._random_trial(run_budget)
            self._log(f"sampled randomly: {nxt.params}")
        else:
            per_budget_models = bool(self.pruner) and not self.interim_results
            if not per_budget_models or max(list(self.models.keys()) + [-np.inf]) <= model_budget:
                self.update_model(model_budget)
            if not self.models:
                nxt = self._random_trial(run_budget)
                self._log(f"sampled randomly: {nxt.params}")
            else:
                if per_budget_models:
                    model_budget = max(self.models.keys())  # the most informed model available
                self._log(f"start sampling routine from model with budget {model_budget}")
                hparams = self.sampling_routine(model_budget)
                nxt = self.create_trial(hparams=hparams, sample_type="model", run_budget=run_budget,
                                        model_budget=model_budget)
                self._log(f"sampled from model with budget {model_budget}: {hparams}")

        retries = 0
        while self.hparams_exist(trial=nxt):
            self._log("sample randomly to encourage exploration")
            nxt = self._random_trial(run_budget, sample_type="random_forced")
            retries += 1
            if retries > 3:
                self._log("not possible to sample new config. Stop Experiment (most/all configs have already been used)")
                return None
        if self.pruner:
            self.pruner.report_trial(original_trial_id=None, new_trial_id=nxt.trial_id)
        self._log(f"start trial {nxt.trial_id}: {nxt.params}, {nxt.info_dict} \n")
        return nxt

    def finalize_experiment(self, trials):
        return

    # ------------------------------------------------------------------ training data for the surrogate
    def include_busy_locations(self) -> bool:
        return self.name() == "GP" and getattr(self, "async_strategy", None) == "impute"

    def _busy_trials(self, budget):
        return [t for t in list(self.trial_store.values())
                if t.info_dict.get("sample_type") == "model" and t.info_dict.get("model_budget") == budget]

    def _require_impute(self):
        if not self.include_busy_locations():
            extra = f"async_strategy: {self.async_strategy}" if self.name() == "GP" else ""
            raise ValueError("Only Optimizer GP with async_strategy == `impute` can include busy locations. "
                             f"Got Optimizer {self.name()} {extra}")

    def get_busy_locations(self, budget=0):
        self._require_impute()
        return np.array([self.searchspace.dict_to_list(t.params) for t in self._busy_trials(budget)], dtype=object)

    def get_imputed_metrics(self, budget=0):
        self._require_impute()
        vals = []
        for t in self._busy_trials(budget):
            v = self.impute_metric(t.params, budget)
            vals.append(v)
            t.info_dict.setdefault("imputed_metrics", []).append(v)
        return np.array(vals, dtype=float)

    def _encode(self, rows):
        names = self.searchspace.keys()
        out = []
        for row in rows:
            row = list(row)[:len(names)] if len(row) > len(names) else list(row)  # drop the trailing budget value
            out.append(self.searchspace.transform(row, normalize_categorical=self.normalize_categorical))
        return np.array(out, dtype=float).reshape(len(out), len(names))

    def _param_rows(self, trials):
        names = self.searchspace.keys()
        return [[t.params[k] for k in names] for t in trials]

    def get_interim_result_idx(self, metric_history, interval=10):
        n = len(metric_history)
        idx = [i for i in range(n) if (i + 1) % interval == 0]
        if not idx or idx[-1] != n - 1:
            idx.append(n - 1)
        return idx

    def get_XY(self, budget=0, interim_results=False, interim_results_interval=10):
        """Training matrix for the surrogate. With ``interim_results`` every kept intermediate metric of a
        trial becomes a row ``[x, n/N]`` (hyperparameters augmented with the normalised budget)."""
        finished = [t for t in self.final_store if self._budget_match(t, budget)]
        sign = -1.0 if self.direction == "max" else 1.0
        X_fin = self._encode(self._param_rows(finished))
        if not interim_results:
            X = X_fin
            y = np.array([sign * t.final_metric for t in finished], dtype=float)
            if self.include_busy_locations():
                busy = self._busy_trials(budget)
                if busy:
                    X = np.concatenate((X, self._encode(self._param_rows(busy))))
                    y = np.concatenate((y, self.get_imputed_metrics(budget)))
        else:
            max_budget = self.get_max_budget()
            rows, ys = [], []
            for x, t in zip(X_fin, finished):
                hist = t.metric_history
                for i in self.get_interim_result_idx(hist, interim_results_interval):
                    frac = float(np.clip(i / max(max_budget - 1, 1), 0.0, 1.0))
                    rows.append(np.append(x, frac))
                    ys.append(sign * hist[i])
            X = np.array(rows, dtype=float).reshape(len(rows), len(self.searchspace.keys()) + 1)
            y = np.array(ys, dtype=float)
            if self.include_busy_locations():
                busy = self._busy_trials(budget)
                if busy:
                    Xb = self._encode(self._param_rows(busy))
                    X = np.concatenate((X, np.append(Xb, np.ones((Xb.shape[0], 1)), 1)))
                    y = np.concatenate((y, self.get_imputed_metrics(budget)))
        assert X.shape[0] == y.shape[0], "X and y need to same first dim"
        return X, y
