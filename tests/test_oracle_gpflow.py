"""Certify the (unpinned) GPflow restatement against closed forms and SciPy."""
import numpy as np
import pytest
from scipy.spatial.distance import cdist

from oracle import cg as ocg
from oracle import gpflow_restated as g
from oracle import models as om


@pytest.mark.parametrize("name", list(g.KERNELS))
def test_kernel_closed_forms(name):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((40, 3))
    Z = rng.standard_normal((25, 3))
    ls = np.array([0.7, 1.3, 2.0])
    k = g.KERNELS[name](variance=1.7, lengthscales=ls)
    r = cdist(X / ls, Z / ls)
    ref = {
        "se": 1.7 * np.exp(-0.5 * r ** 2),
        "matern12": 1.7 * np.exp(-r),
        "matern32": 1.7 * (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r),
        "matern52": 1.7 * (1 + np.sqrt(5) * r + 5 / 3 * r ** 2) * np.exp(-np.sqrt(5) * r),
    }[name]
    np.testing.assert_allclose(k.K(X, Z), ref, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(np.diag(k.K(X)), 1.7, rtol=1e-7)  # Matern: sqrt(1e-36) clamp, not exactly variance
    np.testing.assert_array_equal(k.K_diag(X), np.full(40, 1.7))
    np.testing.assert_allclose(k.K(X), k.K(X, X), rtol=1e-12, atol=1e-13)


def test_square_distance_is_expanded_form_without_clamp():
    x = np.array([[1e8, 1.0]])
    d = g.square_distance(x, x + 1e-9)
    # the expanded form is allowed to go (slightly) negative; GPflow does not clamp it
    assert d.shape == (1, 1)
    X = np.random.default_rng(1).standard_normal((10, 4))
    np.testing.assert_allclose(g.square_distance(X, X), cdist(X, X) ** 2, atol=1e-12)


def test_gaussian_variational_expectations_1d_quadrature():
    lik = g.Gaussian(0.3)
    mu, var, y = np.array([[0.2]]), np.array([[0.5]]), np.array([[1.0]])
    xs, ws = np.polynomial.hermite_e.hermegauss(40)
    f = mu + np.sqrt(var) * xs
    quad = np.sum(ws / np.sqrt(2 * np.pi) * (-0.5 * np.log(2 * np.pi * 0.3) - 0.5 * (y - f) ** 2 / 0.3))
    np.testing.assert_allclose(lik.variational_expectations(None, mu, var, y), [quad], rtol=1e-12)


def test_sgpr_matches_dense_sigma_formulas_and_exact_gp_limit():
    rng = np.random.default_rng(2)
    X = rng.uniform(-2, 2, (60, 2))
    Y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((60, 1))
    k = g.Matern52(variance=1.1, lengthscales=[0.9, 1.4])
    Xs = rng.uniform(-2, 2, (9, 2))
    # Z = X: SGPR == exact GP regression (up to jitter)
    s = g.SGPR((X, Y), k, X.copy(), noise_variance=0.2)
    mean, var = s.predict_f(Xs)
    Kff = k.K(X) + 0.2 * np.eye(60)
    Ksf = k.K(Xs, X)
    np.testing.assert_allclose(mean, Ksf @ np.linalg.solve(Kff, Y), atol=1e-4)
    np.testing.assert_allclose(var[:, 0], k.K_diag(Xs) - np.sum(Ksf * np.linalg.solve(Kff, Ksf.T).T, -1), atol=1e-4)
    lml = -0.5 * (Y.T @ np.linalg.solve(Kff, Y))[0, 0] - 0.5 * np.linalg.slogdet(Kff)[1] - 30 * np.log(2 * np.pi)
    np.testing.assert_allclose(s.elbo(), lml, atol=1e-3)
    # generic Z: predict_f == the Sigma-system formulas the matrix-free path solves by CG
    Z = X[:17].copy()
    s = g.SGPR((X, Y), k, Z, noise_variance=0.2)
    mean, var = s.predict_f(Xs)
    Sigma = s.sigma_matrix()
    kuu = g.Kuu(Z, k, jitter=g.DEFAULT_JITTER)
    exact_sigma = lambda R: np.linalg.solve(Sigma, R)  # noqa: E731
    exact_kuu = lambda R: np.linalg.solve(kuu, R)  # noqa: E731
    m2, v2 = om.sgpr_predict_f_cg(k, X, Y, Z, 0.2, Xs, exact_sigma, exact_kuu)
    np.testing.assert_allclose(m2, mean, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(v2, var, rtol=1e-6, atol=1e-8)
    # matrix-free operator == dense Sigma
    V = rng.standard_normal((3, 17))
    np.testing.assert_allclose(om.sgpr_operator(k, X, Z, 0.2, chunk=16)(V), V @ Sigma, rtol=1e-12)


def test_distances_and_assignment():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((200, 2))
    y = rng.standard_normal((200, 1))
    Z = X[:10].copy()
    k = g.Matern32(variance=1.5, lengthscales=[1.0, 2.0])
    e = om.pairwise_distance(k, "euclidean", X, Z)
    np.testing.assert_allclose(e, cdist(X, Z), atol=1e-13)
    cov = om.pairwise_distance(k, "covariance", X, Z)
    cor = om.pairwise_distance(k, "correlation", X, Z)
    Kxz = k.K(X, Z)
    np.testing.assert_allclose(cov, 3.0 - 2 * Kxz, atol=1e-13)
    np.testing.assert_allclose(cor, 1 - Kxz / 1.5, atol=1e-13)
    # per-point functional form (distance.py as driven by selection.py:24-31)
    for t in ("euclidean", "covariance", "correlation"):
        fn = om.create_distance_fn(k, t)
        np.testing.assert_allclose(fn((Z, X[5])), om.pairwise_distance(k, t, X[5:6], Z)[0], atol=1e-13)
    idx, dist = om.kmeans_indices_and_distances(Z, X)
    np.testing.assert_array_equal(idx, np.argmin(cdist(X, Z), axis=1))
    _, u, counts = om.kmeans_update_inducing_parameters(Z, X, y)
    assert counts.sum() == 200 and counts.shape == (10, 1)
    for j in range(10):
        np.testing.assert_allclose(u[j, 0], y[idx == j].mean())
    _, means, cnt = om.oips_style_assignment(Z, X, y)
    np.testing.assert_array_equal(cnt, counts[:, 0].astype(np.int64))
    np.testing.assert_allclose(means, u[:, 0])


def test_torch_cpu_port_matches_numpy_oracle():
    """The multi-threaded CPU baseline of bench.py is the same algorithm as the NumPy oracle."""
    import torch

    from oracle import torch_cpu as tc

    rng = np.random.default_rng(5)
    X, Z = rng.standard_normal((700, 4)), rng.standard_normal((60, 4))
    ls = np.array([0.9, 1.1, 1.3, 0.7])
    for name in g.KERNELS:
        ok = g.KERNELS[name](variance=1.3, lengthscales=ls)
        Kt = tc.kernel_matrix(name, 1.3, torch.as_tensor(ls), torch.as_tensor(X), torch.as_tensor(Z)).numpy()
        np.testing.assert_allclose(Kt, ok.K(X, Z), rtol=1e-12, atol=1e-14)
    ok = g.Matern52(variance=1.3, lengthscales=ls)
    rhs = rng.standard_normal((2, 60))
    hist = []
    ocg.conjugate_gradient(om.sgpr_operator(ok, X, Z, 0.1), rhs, np.zeros_like(rhs), 0.0, None, 6, 100, history=hist)
    mm = tc.sgpr_operator("matern52", 1.3, torch.as_tensor(ls), torch.as_tensor(X), torch.as_tensor(Z), 0.1, chunk=256)
    _, h = tc.cg_iterations(mm, torch.as_tensor(rhs), 6)
    np.testing.assert_allclose(h.numpy(), np.array(hist), rtol=1e-7)
