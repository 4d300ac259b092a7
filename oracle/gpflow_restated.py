"""NumPy restatement of the GPflow 2.x arithmetic the reference's hot path calls.  TEST INFRASTRUCTURE.

GPflow is a third-party dependency of the reference (``requirements.txt:1``: ``gpflow>=2.5.2``, un-pinned, not
vendored under /root/reference and not installable here), so this file restates its *published* algorithm
(GPflow 2.x ``gpflow/utilities/ops.py::square_distance``, ``gpflow/kernels/stationaries.py``,
``gpflow/covariances/kufs.py|kuus.py``, ``gpflow/likelihoods/scalar_continuous.py::Gaussian``,
``gpflow/models/sgpr.py::SGPR_deprecated``) and anchors on the reference's call sites:
``cggp/models.py:112,141-143,236,255-257,300,333-335`` (Kuu/Kuf/K_diag), ``cggp/models.py:132`` (variational
expectations), ``cggp/optimize.py:50`` (square_distance), ``cggp/cli_utils.py:444-446`` (SGPR),
``cggp/distance.py:17-20,26-29`` (kernel __call__).

PARITY UNPINNED at the last digits: the reference holds no golden vectors for these values.
What pins them: (1) the reference's ONE known-answer test of the kernel values, ``cggp/rff_test.py:9-29`` - the
random-Fourier-feature estimate built by the reference's unmodified ``cggp/rff.py`` (spectral sampling, independent of
these closed forms) must reproduce K(X, X) to rtol 1e-3 / atol 1e-2: fixtures from that file run here
(tests/golden/make_golden_rff.py, 4e6 bases) agree with these kernels to 1e-3 (tests/test_oracle_rff.py), which fixes
the kernel family, the sqrt(3) / sqrt(5) scalings, the variance and the ARD lengthscale convention; (2) closed forms and
the CGGP / ClusterGP / dense cross-checks (tests/test_oracle_gpflow.py); (3) an independent third-party implementation
of the same published formulas - scikit-learn's RBF / Matern kernels and exact GP regression, to 5e-13
(tests/test_oracle_vs_sklearn.py).
"""
from __future__ import annotations

import numpy as np

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter()


def square_distance(X, X2=None):
    """GPflow ``ops.square_distance``: ``|x|^2 + |x2|^2 - 2 x.x2`` (expanded form, NO clamp).

    Call site: ``cggp/optimize.py:50``; used by every stationary kernel below.
    """
    X = np.asarray(X)
    if X2 is None:
        Xs = np.sum(np.square(X), axis=-1, keepdims=True)
        dist = -2.0 * (X @ X.T)
        dist = dist + (Xs + Xs.T)
        return dist
    X2 = np.asarray(X2)
    Xs = np.sum(np.square(X), axis=-1)
    X2s = np.sum(np.square(X2), axis=-1)
    dist = -2.0 * np.tensordot(X, X2, [[-1], [-1]])
    dist = dist + (Xs[..., :, None] + X2s[..., None, :])
    return dist


class Stationary:
    """GPflow ``IsotropicStationary``: ``K = K_r2(square_distance(X/l, X2/l))``."""

    name = "stationary"

    def __init__(self, variance=1.0, lengthscales=1.0, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.variance = self.dtype.type(variance)
        self.lengthscales = np.asarray(lengthscales, dtype=self.dtype)

    def scale(self, X):
        return None if X is None else np.asarray(X, dtype=self.dtype) / self.lengthscales

    def scaled_squared_euclid_dist(self, X, X2=None):
        return square_distance(self.scale(X), self.scale(X2))

    def K_r2(self, r2):
        # GPflow: r = sqrt(maximum(r2, 1e-36)); K_r(r)
        r = np.sqrt(np.maximum(r2, self.dtype.type(1e-36)))
        return self.K_r(r)

    def K_r(self, r):  # pragma: no cover - abstract
        raise NotImplementedError

    def K(self, X, X2=None):
        return self.K_r2(self.scaled_squared_euclid_dist(X, X2))

    def K_diag(self, X):
        X = np.asarray(X)
        return np.full(X.shape[:-1], self.variance, dtype=self.dtype)

    def __call__(self, X, X2=None, *, full_cov=True):
        if not full_cov:
            assert X2 is None
            return self.K_diag(X)
        return self.K(X, X2)


class SquaredExponential(Stationary):
    name = "se"

    def K_r2(self, r2):
        return self.variance * np.exp(self.dtype.type(-0.5) * r2)


class Matern12(Stationary):
    name = "matern12"

    def K_r(self, r):
        return self.variance * np.exp(-r)


class Matern32(Stationary):
    name = "matern32"

    def K_r(self, r):
        sqrt3 = self.dtype.type(np.sqrt(3.0))
        return self.variance * (self.dtype.type(1.0) + sqrt3 * r) * np.exp(-sqrt3 * r)


class Matern52(Stationary):
    name = "matern52"

    def K_r(self, r):
        sqrt5 = self.dtype.type(np.sqrt(5.0))
        one = self.dtype.type(1.0)
        c = self.dtype.type(5.0 / 3.0)
        return self.variance * (one + sqrt5 * r + c * np.square(r)) * np.exp(-sqrt5 * r)


KERNELS = {k.name: k for k in (SquaredExponential, Matern12, Matern32, Matern52)}


def Kuu(Z, kernel, jitter=0.0):
    """GPflow ``covariances.Kuu(InducingPoints)``: ``K(Z) + jitter*I`` (``cggp/models.py:300,333``)."""
    Kzz = kernel.K(Z)
    Kzz = Kzz + jitter * np.eye(Kzz.shape[0], dtype=Kzz.dtype)
    return Kzz


def Kuf(Z, kernel, Xnew):
    """GPflow ``covariances.Kuf(InducingPoints)``: ``K(Z, Xnew)`` [M, N] (``cggp/models.py:334``)."""
    return kernel.K(Z, Xnew)


class Gaussian:
    """GPflow ``likelihoods.Gaussian`` (variance only)."""

    def __init__(self, variance=1.0, dtype=np.float64):
        self.variance = np.dtype(dtype).type(variance)

    def variational_expectations(self, X, Fmu, Fvar, Y):
        """``-0.5 log 2pi - 0.5 log s2 - 0.5 ((Y-Fmu)^2 + Fvar)/s2`` summed over the last axis -> [N]."""
        v = self.variance
        ve = -0.5 * np.log(2.0 * np.pi) - 0.5 * np.log(v) - 0.5 * (np.square(Y - Fmu) + Fvar) / v
        return np.sum(ve, axis=-1)

    def predict_log_density(self, X, Fmu, Fvar, Y):
        s2 = Fvar + self.variance
        ld = -0.5 * (np.log(2.0 * np.pi) + np.log(s2) + np.square(Y - Fmu) / s2)
        return np.sum(ld, axis=-1)


class SGPR:
    """GPflow ``models.SGPR`` (Titsias 2009), zero mean function; built by ``cggp/cli_utils.py:444-446``."""

    def __init__(self, data, kernel, Z, noise_variance=1.0, jitter=DEFAULT_JITTER):
        self.X, self.Y = (np.asarray(d) for d in data)
        self.kernel = kernel
        self.Z = np.asarray(Z)
        self.noise_variance = noise_variance
        self.jitter = jitter

    def _common(self):
        from scipy.linalg import cholesky, solve_triangular

        sigma_sq = self.noise_variance
        sigma = np.sqrt(sigma_sq)
        kuf = Kuf(self.Z, self.kernel, self.X)
        kuu = Kuu(self.Z, self.kernel, jitter=self.jitter)
        L = cholesky(kuu, lower=True)
        A = solve_triangular(L, kuf, lower=True) / sigma
        AAT = A @ A.T
        B = AAT + np.eye(AAT.shape[0], dtype=AAT.dtype)
        LB = cholesky(B, lower=True)
        return A, AAT, L, LB, sigma, sigma_sq

    def elbo(self):
        from scipy.linalg import solve_triangular

        A, AAT, L, LB, sigma, sigma_sq = self._common()
        N = self.X.shape[0]
        P = self.Y.shape[1]
        const = -0.5 * N * P * np.log(2.0 * np.pi)
        half_logdet_B = np.sum(np.log(np.diag(LB)))
        log_sigma_sq = N * np.log(sigma_sq)
        trace_k = np.sum(self.kernel.K_diag(self.X) / sigma_sq)
        trace_q = np.sum(np.diag(AAT))
        logdet = -P * (half_logdet_B + 0.5 * log_sigma_sq + 0.5 * (trace_k - trace_q))
        err = self.Y / sigma
        Aerr = A @ err
        c = solve_triangular(LB, Aerr, lower=True)
        quad = -0.5 * (np.sum(np.square(err)) - np.sum(np.square(c)))
        return const + logdet + quad

    def predict_f(self, Xnew):
        from scipy.linalg import solve_triangular

        A, AAT, L, LB, sigma, sigma_sq = self._common()
        Kus = Kuf(self.Z, self.kernel, Xnew)
        Aerr = A @ self.Y
        c = solve_triangular(LB, Aerr, lower=True) / sigma
        tmp1 = solve_triangular(L, Kus, lower=True)
        tmp2 = solve_triangular(LB, tmp1, lower=True)
        mean = tmp2.T @ c
        var = (
            self.kernel.K_diag(Xnew)
            + np.sum(np.square(tmp2), axis=0)
            - np.sum(np.square(tmp1), axis=0)
        )
        var = np.tile(var[:, None], [1, self.Y.shape[1]])
        return mean, var

    # The same posterior through the linear system Sigma = Kuu + s^-2 Kuf Kfu = L B L^T that GPflow factorises.
    def sigma_matrix(self):
        kuf = Kuf(self.Z, self.kernel, self.X)
        kuu = Kuu(self.Z, self.kernel, jitter=self.jitter)
        return kuu + (kuf @ kuf.T) / self.noise_variance
