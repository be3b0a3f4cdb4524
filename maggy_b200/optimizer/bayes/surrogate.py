"""Numpy/scipy surrogate models for the Bayesian optimizers.

The reference leans on scikit-optimize's ``GaussianProcessRegressor`` (Matern-5/2 x constant kernel,
Gaussian noise, normalised targets; maggy/optimizer/bayes/gp.py:266-291) and statsmodels'
``KDEMultivariate`` (maggy/optimizer/bayes/tpe.py:182-189). Neither library exists in this image, and
the driver stays on the CPU by design, so both are implemented here from first principles:

* :class:`GaussianProcess` — ARD Matern-5/2 with amplitude and a learned white-noise term, targets
  standardised, hyperparameters by L-BFGS-B on the exact log marginal likelihood (analytic gradient,
  random restarts), predictions with analytic input gradients (needed by the acquisition optimiser).
* :class:`ProductKDE` — product kernel density estimate: Gaussian kernels for continuous dimensions,
  Aitchison-Aitken kernels for unordered categorical ones, normal-reference bandwidths.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_solve, solve_triangular
from scipy.optimize import fmin_l_bfgs_b

_SQRT5 = np.sqrt(5.0)


def _matern52(X1, X2, ls):
    """k(r) = (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r), r = |(x1-x2)/ls|. Returns K and r."""
    d = (X1[:, None, :] - X2[None, :, :]) / ls
    r = np.sqrt(np.maximum((d * d).sum(-1), 0.0))
    s = _SQRT5 * r
    return (1.0 + s + s * s / 3.0) * np.exp(-s), r


class GaussianProcess:
    """Exact GP regression. theta = [log amp, log ls_1..d, log noise]."""

    def __init__(self, n_dims, amp_bounds=(0.01, 1000.0), ls_bounds=(0.01, 100.0), noise_bounds=(1e-5, 1e5),
                 n_restarts_optimizer=2, normalize_y=True, random_state=None):
        self.n_dims = n_dims
        self.bounds = np.log(np.array([amp_bounds] + [ls_bounds] * n_dims + [noise_bounds]))
        self.n_restarts = n_restarts_optimizer
        self.normalize_y = normalize_y
        self.rng = np.random.RandomState(random_state)
        self.theta = np.zeros(n_dims + 2)
        self.theta[-1] = np.log(1.0)
        self.X = None

    def clone(self):
        return GaussianProcess(self.n_dims, n_restarts_optimizer=self.n_restarts, normalize_y=self.normalize_y,
                               random_state=self.rng.randint(1 << 30))

    # ------------------------------------------------------------------ marginal likelihood
    def _nlml(self, theta, X, y):
        amp, ls, noise = np.exp(theta[0]), np.exp(theta[1:-1]), np.exp(theta[-1])
        n = X.shape[0]
        Km, r = _matern52(X, X, ls)
        K = amp * Km + (noise + 1e-10) * np.eye(n)
        try:
            L = np.linalg.cholesky(K)
        except np.linalg.LinAlgError:
            return 1e25, np.zeros_like(theta)
        alpha = cho_solve((L, True), y)
        nlml = 0.5 * y @ alpha + np.log(np.diag(L)).sum() + 0.5 * n * np.log(2 * np.pi)
        # gradient: dNLML/dtheta_j = -0.5 tr((alpha alpha^T - K^-1) dK/dtheta_j)
        Kinv = cho_solve((L, True), np.eye(n))
        W = np.outer(alpha, alpha) - Kinv
        grad = np.empty_like(theta)
        grad[0] = -0.5 * np.sum(W * (amp * Km))
        # d k / d log ls_j = amp * (5/3)(1 + sqrt5 r) exp(-sqrt5 r) * (dx_j/ls_j)^2
        base = amp * (5.0 / 3.0) * (1.0 + _SQRT5 * r) * np.exp(-_SQRT5 * r)
        for j in range(self.n_dims):
            dj = (X[:, None, j] - X[None, :, j]) / ls[j]
            grad[1 + j] = -0.5 * np.sum(W * (base * dj * dj))
        grad[-1] = -0.5 * np.trace(W) * noise
        return nlml, grad

    def fit(self, X, y):
        X = np.asarray(X, dtype=float)
        y = np.asarray(y, dtype=float).ravel()
        if self.normalize_y:
            self.y_mean = y.mean()
            self.y_std = y.std() if y.std() > 0 else 1.0
        else:
            self.y_mean, self.y_std = 0.0, 1.0
        yn = (y - self.y_mean) / self.y_std
        starts = [np.clip(self.theta, self.bounds[:, 0], self.bounds[:, 1])]
        for _ in range(self.n_restarts):
            starts.append(self.rng.uniform(self.bounds[:, 0], self.bounds[:, 1]))
        best = (np.inf, starts[0])
        for t0 in starts:
            try:
                t, f, _ = fmin_l_bfgs_b(self._nlml, t0, args=(X, yn), bounds=list(map(tuple, self.bounds)), maxiter=200)
            except Exception:
                continue
            if f < best[0]:
                best = (f, t)
        self.theta = best[1]
        self.X = X
        amp, ls, noise = np.exp(self.theta[0]), np.exp(self.theta[1:-1]), np.exp(self.theta[-1])
        Km, _ = _matern52(X, X, ls)
        K = amp * Km + (noise + 1e-10) * np.eye(X.shape[0])
        self.L = np.linalg.cholesky(K)
        self.alpha = cho_solve((self.L, True), yn)
        self.log_marginal_likelihood_value_ = -best[0]
        return self

    # ------------------------------------------------------------------ prediction
    def _kstar(self, Xs):
        amp, ls = np.exp(self.theta[0]), np.exp(self.theta[1:-1])
        Km, r = _matern52(Xs, self.X, ls)
        return amp * Km, r, amp, ls

    def predict(self, Xs, return_std=False, return_mean_grad=False, return_std_grad=False):
        """Posterior mean (and std) of the latent function at Xs. Gradients are wrt a single point."""
        Xs = np.atleast_2d(np.asarray(Xs, dtype=float))
        Ks, r, amp, ls = self._kstar(Xs)
        mean = Ks @ self.alpha * self.y_std + self.y_mean
        out = [mean]
        if return_std or return_std_grad:
            v = solve_triangular(self.L, Ks.T, lower=True)
            var = np.maximum(amp - (v * v).sum(0), 1e-12)
            std = np.sqrt(var) * self.y_std
            out.append(std)
        if return_mean_grad or return_std_grad:
            assert Xs.shape[0] == 1, "gradients are computed for one point at a time"
            # dk(x*, xi)/dx*_j = -amp (5/3)(1 + sqrt5 r) exp(-sqrt5 r) (x*_j - xi_j)/ls_j^2
            coef = -amp * (5.0 / 3.0) * (1.0 + _SQRT5 * r[0]) * np.exp(-_SQRT5 * r[0])  # [n]
            dK = coef[:, None] * (Xs[0][None, :] - self.X) / (ls * ls)[None, :]          # [n, d]
            mean_grad = dK.T @ self.alpha * self.y_std
            out.append(mean_grad)
            if return_std_grad:
                Kinv_ks = cho_solve((self.L, True), Ks[0])
                var_grad = -2.0 * dK.T @ Kinv_ks
                std_grad = var_grad * self.y_std * self.y_std / (2.0 * max(out[1][0], 1e-12))
                out.append(std_grad)
        return out[0] if len(out) == 1 else tuple(out)

    def sample_y(self, Xs, n_samples=1, random_state=None):
        """Draw joint posterior samples at Xs: array [len(Xs), n_samples]."""
        Xs = np.atleast_2d(np.asarray(Xs, dtype=float))
        Ks, _, amp, ls = self._kstar(Xs)
        Kss, _ = _matern52(Xs, Xs, ls)
        v = solve_triangular(self.L, Ks.T, lower=True)
        cov = amp * Kss - v.T @ v
        cov = 0.5 * (cov + cov.T) + 1e-8 * np.eye(Xs.shape[0])
        mean = Ks @ self.alpha
        rng = np.random.RandomState(random_state) if random_state is not None else self.rng
        try:
            Lc = np.linalg.cholesky(cov)
            z = rng.standard_normal((Xs.shape[0], n_samples))
            s = mean[:, None] + Lc @ z
        except np.linalg.LinAlgError:
            s = rng.multivariate_normal(mean, cov, n_samples, check_valid="ignore").T
        return s * self.y_std + self.y_mean


class ProductKDE:
    """Product-kernel density estimate over mixed continuous ('c') / unordered categorical ('u') data."""

    def __init__(self, data, var_type: str, n_categories=None, bw="normal_reference"):
        self.data = np.atleast_2d(np.asarray(data, dtype=float))
        self.var_type = var_type
        n, d = self.data.shape
        assert len(var_type) == d
        self.n_categories = list(n_categories) if n_categories is not None else [
            int(self.data[:, j].max()) + 1 if var_type[j] == "u" else 0 for j in range(d)]
        if isinstance(bw, str):
            if bw != "normal_reference":
                raise ValueError(f"unknown bandwidth rule {bw}")
            # Scott/Silverman normal reference rule, applied per column like statsmodels does
            sigma = self.data.std(axis=0, ddof=1) if n > 1 else np.ones(d)
            self.bw = 1.06 * sigma * n ** (-1.0 / (4.0 + d))
        else:
            self.bw = np.asarray(bw, dtype=float)
        # categorical bandwidths are probabilities of leaving the observed category
        for j in range(d):
            if var_type[j] == "u":
                c = max(self.n_categories[j], 2)
                self.bw[j] = float(np.clip(self.bw[j], 0.0, (c - 1.0) / c))

    def pdf(self, x) -> float:
        x = np.asarray(x, dtype=float).ravel()
        dens = np.ones(self.data.shape[0])
        for j, kind in enumerate(self.var_type):
            col = self.data[:, j]
            h = self.bw[j]
            if kind == "c":
                h = max(h, 1e-10)
                z = (x[j] - col) / h
                dens *= np.exp(-0.5 * z * z) / (np.sqrt(2 * np.pi) * h)
            else:
                c = max(self.n_categories[j], 2)
                dens *= np.where(col == x[j], 1.0 - h, h / (c - 1.0))
        return float(dens.mean())
