"""Acquisition functions on a fitted GP surrogate (reference: maggy/optimizer/bayes/acquisitions.py:25-193,
which wraps scikit-optimize's ``_gaussian_acquisition``). All are *minimised*: lower value = more
promising point. ``evaluate`` scores a batch; ``evaluate_1_d`` returns ``(value, gradient)`` for one
point so L-BFGS-B can polish the best candidates."""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np
from scipy.stats import norm


class AbstractAcquisitionFunction(ABC):
    @staticmethod
    @abstractmethod
    def evaluate(X, surrogate_model, y_opt, acq_func_kwargs=None):
        ...

    @staticmethod
    @abstractmethod
    def evaluate_1_d(x, surrogate_model, y_opt, acq_func_kwargs=None):
        ...

    def name(self):
        return type(self).__name__


def _kw(kwargs, key, default):
    return default if not kwargs else kwargs.get(key, default)


class GaussianProcess_EI(AbstractAcquisitionFunction):
    """Negative expected improvement over ``y_opt`` (minimisation convention), exploration margin ``xi``."""

    @staticmethod
    def evaluate(X, surrogate_model, y_opt, acq_func_kwargs=None):
        xi = _kw(acq_func_kwargs, "xi", 0.01)
        mu, std = surrogate_model.predict(X, return_std=True)
        imp = y_opt - xi - mu
        z = imp / std
        return -(imp * norm.cdf(z) + std * norm.pdf(z))

    @staticmethod
    def evaluate_1_d(x, surrogate_model, y_opt, acq_func_kwargs=None):
        xi = _kw(acq_func_kwargs, "xi", 0.01)
        mu, std, dmu, dstd = surrogate_model.predict(np.asarray(x).reshape(1, -1), return_std=True,
                                                      return_mean_grad=True, return_std_grad=True)
        mu, std = mu[0], std[0]
        imp = y_opt - xi - mu
        z = imp / std
        cdf, pdf = norm.cdf(z), norm.pdf(z)
        ei = imp * cdf + std * pdf
        grad = -dmu * cdf + dstd * pdf  # d EI / dx
        return -ei, -grad


class GaussianProcess_PI(AbstractAcquisitionFunction):
    """Negative probability of improvement."""

    @staticmethod
    def evaluate(X, surrogate_model, y_opt, acq_func_kwargs=None):
        xi = _kw(acq_func_kwargs, "xi", 0.01)
        mu, std = surrogate_model.predict(X, return_std=True)
        return -norm.cdf((y_opt - xi - mu) / std)

    @staticmethod
    def evaluate_1_d(x, surrogate_model, y_opt, acq_func_kwargs=None):
        xi = _kw(acq_func_kwargs, "xi", 0.01)
        mu, std, dmu, dstd = surrogate_model.predict(np.asarray(x).reshape(1, -1), return_std=True,
                                                      return_mean_grad=True, return_std_grad=True)
        mu, std = mu[0], std[0]
        z = (y_opt - xi - mu) / std
        grad = norm.pdf(z) * (-dmu / std - z * dstd / std)
        return -norm.cdf(z), -grad


class GaussianProcess_LCB(AbstractAcquisitionFunction):
    """Lower confidence bound ``mu - kappa * std``."""

    @staticmethod
    def evaluate(X, surrogate_model, y_opt, acq_func_kwargs=None):
        kappa = _kw(acq_func_kwargs, "kappa", 1.96)
        mu, std = surrogate_model.predict(X, return_std=True)
        return mu - kappa * std

    @staticmethod
    def evaluate_1_d(x, surrogate_model, y_opt, acq_func_kwargs=None):
        kappa = _kw(acq_func_kwargs, "kappa", 1.96)
        mu, std, dmu, dstd = surrogate_model.predict(np.asarray(x).reshape(1, -1), return_std=True,
                                                      return_mean_grad=True, return_std_grad=True)
        return mu[0] - kappa * std[0], dmu - kappa * dstd


class AsyTS(AbstractAcquisitionFunction):
    """Asynchronous Thompson sampling: one joint posterior draw over the candidate set."""

    @staticmethod
    def evaluate(X, surrogate_model, y_opt, acq_func_kwargs=None):
        return surrogate_model.sample_y(X, n_samples=1).ravel()

    @staticmethod
    def evaluate_1_d(x, surrogate_model, y_opt, acq_func_kwargs=None):
        return float(surrogate_model.sample_y(np.asarray(x).reshape(1, -1), n_samples=1).ravel()[0])


class GaussianProcess_UCB(AbstractAcquisitionFunction):
    @staticmethod
    def evaluate(X, surrogate_model, y_opt, acq_func_kwargs=None):
        raise NotImplementedError("UCB is not available: the optimizers minimise, use LCB")

    evaluate_1_d = evaluate


class TPE_EI(AbstractAcquisitionFunction):
    @staticmethod
    def evaluate(X, surrogate_model, y_opt, acq_func_kwargs=None):
        raise NotImplementedError("TPE computes its density ratio inside maggy_b200.optimizer.bayes.TPE")

    evaluate_1_d = evaluate


class HLP(AbstractAcquisitionFunction):
    @staticmethod
    def evaluate(X, surrogate_model, y_opt, acq_func_kwargs=None):
        raise NotImplementedError("hallucination with local penalisation is not implemented")

    evaluate_1_d = evaluate
