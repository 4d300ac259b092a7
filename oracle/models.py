"""NumPy restatement of ``cggp/models.py`` (ClusterGP / CGGP), ``cggp/utils.py::add_diagonal``,
``cggp/distance.py`` and the nearest-centre assignment of ``cggp/optimize.py:41-98`` /
``cggp/selection.py:14-32``.  TEST INFRASTRUCTURE, not product code.

Model formulas are pinned against the reference's unmodified ``cggp/models.py`` run over a NumPy shim
(``tests/golden/make_golden.py``); the kernel arithmetic underneath is the unpinned GPflow restatement
(``oracle/gpflow_restated.py``).
"""
from __future__ import annotations

import numpy as np

from . import gpflow_restated as gpf
from .cg import ConjugateGradient


def add_diagonal(matrix, diagonal):
    """cggp/utils.py:11-17."""
    out = np.array(matrix, copy=True)
    idx = np.arange(out.shape[0])
    out[idx, idx] = out[idx, idx] + diagonal
    return out


class ClusterGP:
    """cggp/models.py:176-276 (Cholesky comparator), zero mean function."""

    def __init__(self, kernel, likelihood, Z, *, cluster_counts=None, num_data=None, pseudo_u=None):
        self.kernel = kernel
        self.likelihood = likelihood
        self.Z = np.asarray(Z)
        m = self.Z.shape[0]
        dtype = self.Z.dtype
        self.pseudo_u = np.zeros((m, 1), dtype) if pseudo_u is None else np.asarray(pseudo_u, dtype)
        if self.pseudo_u.shape != (m, 1):
            raise ValueError("Pseudo-u argument shape must match actual pseudo-u shape.")  # :203-206
        self.cluster_counts = np.ones((m, 1), dtype) if cluster_counts is None else np.asarray(cluster_counts, dtype)
        if self.cluster_counts.shape != (m, 1):
            raise ValueError("Cluster counts argument shape must match pseudo-u shape.")  # :208-210
        self.num_data = num_data

    @property
    def diag_variance(self):  # :226-228
        return self.likelihood.variance / self.cluster_counts

    def scale(self, batch_size, dtype):  # :163-169
        if self.num_data is not None:
            return dtype.type(self.num_data) / dtype.type(batch_size)
        return dtype.type(1.0)

    def _Kmm_KmmLambda(self):
        Kmm = gpf.Kuu(self.Z, self.kernel, jitter=0.0)
        return Kmm, add_diagonal(Kmm, self.diag_variance[:, 0])

    def prior_kl(self):  # :230-248
        from scipy.linalg import cho_solve, cholesky

        Kmm, K = self._Kmm_KmmLambda()
        var = self.diag_variance
        L = cholesky(K, lower=True)
        a = cho_solve((L, True), self.pseudo_u)
        quad = np.sum((Kmm @ a) * a)
        trace = np.trace(cho_solve((L, True), Kmm))
        logdet = np.sum(2.0 * np.log(np.diag(L)))
        const = np.sum(np.log(var))
        return 0.5 * (quad - trace + logdet - const)

    def predict_f(self, Xnew, full_cov=False):  # :250-276
        from scipy.linalg import cho_solve, cholesky, solve_triangular

        Kmm, K = self._Kmm_KmmLambda()
        Kmn = gpf.Kuf(self.Z, self.kernel, Xnew)
        Knn = self.kernel.K(Xnew) if full_cov else self.kernel.K_diag(Xnew)
        L = cholesky(K, lower=True)
        a = cho_solve((L, True), self.pseudo_u)
        A = solve_triangular(L, Kmn, lower=True)
        if not full_cov:
            fvar = (Knn - np.sum(np.square(A), axis=0))[:, None]
        else:
            fvar = (Knn - A.T @ A)[None, ...]
        fmu = Kmn.T @ a
        return fmu, fvar

    def elbo(self, data):  # LpSVGP.elbo, :125-134
        x, y = data
        kl = self.prior_kl()
        f_mean, f_var = self.predict_f(x)
        var_exp = self.likelihood.variational_expectations(x, f_mean, f_var, y)
        scale = self.scale(x.shape[0], np.asarray(kl).dtype)
        return np.sum(var_exp) * scale - kl


class CGGP(ClusterGP):
    """cggp/models.py:279-354 ("CDGP"): every solve by CG; ``logdet`` forward value is 0 (:46, :319)."""

    def __init__(self, kernel, likelihood, Z, conjugate_gradient: ConjugateGradient, num_probes=5, **kw):
        super().__init__(kernel, likelihood, Z, **kw)
        self.conjugate_gradient = conjugate_gradient
        self.num_probes = num_probes
        self.probes = None  # injected Rademacher probes [M, P] (the reference draws from the global TF RNG, :310)

    def prior_kl(self):  # :293-322
        Kmm, KmmLambda = self._Kmm_KmmLambda()
        var = self.diag_variance
        cg = self.conjugate_gradient
        a = cg(KmmLambda, self.pseudo_u)  # :303
        if self.num_probes is None:  # :304-306
            trace = np.trace(cg(KmmLambda, Kmm))
        else:  # :308-314
            probes = self.probes
            assert probes is not None and probes.shape == (Kmm.shape[0], self.num_probes)
            sol = cg(KmmLambda, probes)
            trace = np.sum(sol * (Kmm @ probes)) / Kmm.dtype.type(self.num_probes)
        quad = np.sum((Kmm @ a) * a)  # :316-317
        logdet = Kmm.dtype.type(0.0)  # eval_logdet forward value, :46
        const = np.sum(np.log(var))  # :321
        return 0.5 * (quad - trace + logdet - const)

    def predict_f(self, Xnew, full_cov=False):  # :324-354
        Kmm, KmmLambda = self._Kmm_KmmLambda()
        Kmn = gpf.Kuf(self.Z, self.kernel, Xnew)
        Knn = self.kernel.K(Xnew) if full_cov else self.kernel.K_diag(Xnew)
        cg = self.conjugate_gradient
        a = cg(KmmLambda, self.pseudo_u)  # :339
        S = cg(KmmLambda, Kmn)  # :340
        if not full_cov:
            fvar = (Knn - np.sum(Kmn * S, axis=0))[:, None]  # :343-345
        else:
            fvar = (Knn - Kmn.T @ S)[None, ...]  # :347-349
        fmu = Kmn.T @ a  # :351
        return fmu, fvar


def eval_logdet_grad(matrix, cg, df=1.0, probes=None):
    """Gradient of cggp/models.py::eval_logdet (:30-44): ``df * (A^-1)^T`` or the Hutchinson estimate."""
    n = matrix.shape[-1]
    if probes is None:
        return df * cg(matrix, np.eye(n, dtype=matrix.dtype)).T
    lv = cg(matrix, probes)
    return (lv @ (df * probes).T) / matrix.dtype.type(probes.shape[1])


# ---------------------------------------------------------------------------------------------------------
# The north-star operator Sigma = Kuu + s^-2 Kuf Kfu (SGPR's system), applied matrix-free.
# ---------------------------------------------------------------------------------------------------------
def kuf_kfu_matmul(kernel, X, Z, V, chunk=8192):
    """``V @ (Kuf Kfu)`` for row-vectors V [B, M] without materialising Kfu [N, M]: chunked over N."""
    X = np.asarray(X)
    out = np.zeros_like(V)
    for s in range(0, X.shape[0], chunk):
        Kfu = kernel.K(X[s:s + chunk], Z)  # [c, M]
        T = V @ Kfu.T  # [B, c]
        out += T @ Kfu
    return out


def sgpr_operator(kernel, X, Z, noise_variance, jitter=gpf.DEFAULT_JITTER, chunk=8192):
    """Callable ``V -> V @ (Kuu + jitter I + s^-2 Kuf Kfu)`` for ``oracle.cg.conjugate_gradient``."""
    kuu = gpf.Kuu(Z, kernel, jitter=jitter)

    def matmul(V):
        return V @ kuu + kuf_kfu_matmul(kernel, X, Z, V, chunk) / noise_variance

    return matmul


def sgpr_predict_f_cg(kernel, X, Y, Z, noise_variance, Xnew, cg_sigma, cg_kuu, jitter=gpf.DEFAULT_JITTER):
    """SGPR ``predict_f`` (diag) written on the linear system GPflow factorises:
    ``mean = Ksu Sigma^-1 Kuf y / s2``; ``var = k** - Ksu Kuu^-1 Kus + Ksu Sigma^-1 Kus``.
    ``cg_sigma(rhs [M,B])`` / ``cg_kuu(rhs)`` are solver callables (CG or exact).
    """
    Kus = gpf.Kuf(Z, kernel, Xnew)  # [M, B]
    kfu_y = np.zeros((Z.shape[0], Y.shape[1]), dtype=Y.dtype)
    for s in range(0, X.shape[0], 8192):
        kfu_y += kernel.K(Z, X[s:s + 8192]) @ Y[s:s + 8192]
    c = cg_sigma(kfu_y / noise_variance)  # [M, P]
    mean = Kus.T @ c
    S1 = cg_kuu(Kus)
    S2 = cg_sigma(Kus)
    var = kernel.K_diag(Xnew) - np.sum(Kus * S1, axis=0) + np.sum(Kus * S2, axis=0)
    return mean, np.tile(var[:, None], [1, Y.shape[1]])


# ---------------------------------------------------------------------------------------------------------
# cggp/distance.py and the assignment step that consumes it
# ---------------------------------------------------------------------------------------------------------
def euclid_distance(args):
    """cggp/distance.py:9-11 (difference form, then 2-norm over the last axis)."""
    x, y = args
    return np.linalg.norm(np.asarray(x) - np.asarray(y), axis=-1)


def create_distance_fn(kernel, distance_type):
    """cggp/distance.py:14-34; ``(x [M,D], y [D]) -> [M]`` as driven by selection.py:24-31."""

    def _kxy(x, y):
        x = np.atleast_2d(x)
        y = np.atleast_2d(y)
        return kernel.K(x, y)

    def cov(args):  # :15-22
        x, y = args
        x_dist = kernel.K_diag(np.atleast_2d(x))
        y_dist = kernel.K_diag(np.atleast_2d(y))
        xy = _kxy(x, y)
        return np.squeeze(x_dist[:, None] + y_dist[None, :] - 2 * xy)

    def cor(args):  # :24-30
        x, y = args
        x_dist = kernel.K_diag(np.atleast_2d(x))
        y_dist = kernel.K_diag(np.atleast_2d(y))
        xy = _kxy(x, y)
        return np.squeeze(1.0 - xy / np.sqrt(x_dist[:, None] * y_dist[None, :]))

    return {"covariance": cov, "correlation": cor, "euclidean": euclid_distance}[distance_type]


def pairwise_distance(kernel, distance_type, X, Z):
    """[N, M] matrix of ``distance_fn((Z, x_i))`` rows (what selection.py:24-31 evaluates under vectorized_map)."""
    X = np.asarray(X)
    Z = np.asarray(Z)
    if distance_type == "euclidean":
        return np.linalg.norm(X[:, None, :] - Z[None, :, :], axis=-1)
    Kxz = kernel.K(X, Z)
    kx = kernel.K_diag(X)[:, None]
    kz = kernel.K_diag(Z)[None, :]
    if distance_type == "covariance":
        return kx + kz - 2 * Kxz
    if distance_type == "correlation":
        return 1.0 - Kxz / np.sqrt(kx * kz)
    raise ValueError(distance_type)


def kmeans_indices_and_distances(centroids, points, kernel=None, distance_type="euclidean", chunk=4096):
    """cggp/selection.py:14-32: argmin over centroids per point, and the chosen distance."""
    idx = np.empty(points.shape[0], dtype=np.int64)
    dist = np.empty(points.shape[0], dtype=points.dtype)
    for s in range(0, points.shape[0], chunk):
        Dm = pairwise_distance(kernel, distance_type, points[s:s + chunk], centroids)
        ii = np.argmin(Dm, axis=-1)
        idx[s:s + chunk] = ii
        dist[s:s + chunk] = Dm[np.arange(Dm.shape[0]), ii]
    return idx, dist


def kmeans_update_inducing_parameters(Z, x, y, kernel=None, distance_type="euclidean"):
    """cggp/optimize.py:81-98: ``(Z, u = cluster means of y, counts)``; empty clusters give 0/0 = nan as there."""
    m = Z.shape[0]
    idx, _ = kmeans_indices_and_distances(Z, x, kernel, distance_type)
    counts = np.bincount(idx, minlength=m).astype(Z.dtype)[:, None]
    u = np.zeros((m, 1), dtype=Z.dtype)
    np.add.at(u, (idx, 0), y[:, 0])
    with np.errstate(divide="ignore", invalid="ignore"):
        u = u / counts
    return Z, u, counts


def oips_style_assignment(Z, x, y):
    """cggp/optimize.py:50-78: assignment by ``argmin(square_distance(iv, inputs), axis=0)``, per-cluster mean of y
    (nan for empty clusters, as ``tf.reduce_mean`` of an empty tensor) and counts with empties replaced by 1."""
    m = Z.shape[0]
    d = gpf.square_distance(Z, x)  # [M, N]
    idx = np.argmin(d, axis=0)
    counts = np.bincount(idx, minlength=m).astype(np.int64)
    sums = np.zeros(m, dtype=y.dtype)
    np.add.at(sums, idx, y[:, 0])
    with np.errstate(divide="ignore", invalid="ignore"):
        means = sums / counts
    new_counts = np.where(counts != 0, counts, 1)
    return Z, means, new_counts
