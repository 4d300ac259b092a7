"""oracle.models vs golden vectors produced by the reference's own cggp/models.py (tests/golden)."""
import numpy as np
import pytest

from oracle import cg as ocg
from oracle import gpflow_restated as g
from oracle import models as om

CASES = ["se_exacttrace", "matern32_probes", "matern52_exacttrace"]


def build(c):
    kernel = g.KERNELS[str(c["kernel"])](variance=float(c["variance"]), lengthscales=c["lengthscales"])
    lik = g.Gaussian(variance=float(c["noise"]))
    probes = c["probes"]
    num_probes = None if probes.shape[1] == 0 else probes.shape[1]
    cg = ocg.ConjugateGradient(float(c["thr"]))
    m = om.CGGP(kernel, lik, c["Z"], cg, num_probes=num_probes, cluster_counts=c["counts"], pseudo_u=c["u"],
                num_data=int(c["num_data"]))
    m.probes = None if num_probes is None else probes
    cl = om.ClusterGP(kernel, lik, c["Z"], cluster_counts=c["counts"], pseudo_u=c["u"], num_data=int(c["num_data"]))
    return m, cl


@pytest.mark.parametrize("name", CASES)
def test_cggp_matches_reference_run(models_golden, name):
    c = models_golden[name]
    m, cl = build(c)
    np.testing.assert_allclose(m.prior_kl(), c["cggp_kl"], rtol=1e-9)
    mu, var = m.predict_f(c["Xnew"])
    np.testing.assert_allclose(mu, c["cggp_mu"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, c["cggp_var"], rtol=1e-8, atol=1e-10)
    _, var_fc = m.predict_f(c["Xnew"], full_cov=True)
    np.testing.assert_allclose(var_fc, c["cggp_var_fullcov"], rtol=1e-8, atol=1e-10)
    elbo = m.elbo((c["X"][:200], c["y"][:200]))
    np.testing.assert_allclose(elbo, c["cggp_elbo"], rtol=1e-9)


@pytest.mark.parametrize("name", CASES)
def test_clustergp_matches_reference_run(models_golden, name):
    c = models_golden[name]
    _, cl = build(c)
    np.testing.assert_allclose(cl.prior_kl(), c["cluster_kl"], rtol=1e-9)
    mu, var = cl.predict_f(c["Xnew"])
    np.testing.assert_allclose(mu, c["cluster_mu"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, c["cluster_var"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(cl.elbo((c["X"][:200], c["y"][:200])), c["cluster_elbo"], rtol=1e-9)


@pytest.mark.parametrize("name", ["se_exacttrace", "matern52_exacttrace"])
def test_cggp_close_to_cholesky_comparator(models_golden, name):
    """CGGP == ClusterGP in exact arithmetic apart from logdet := 0 (models.py:46,319 vs :245)."""
    c = models_golden[name]
    m, cl = build(c)
    mu, var = m.predict_f(c["Xnew"])
    cmu, cvar = cl.predict_f(c["Xnew"])
    np.testing.assert_allclose(mu, cmu, atol=5e-3)
    np.testing.assert_allclose(var, cvar, atol=5e-3)
    Kmm, KmmLambda = cl._Kmm_KmmLambda()
    logdet = np.linalg.slogdet(KmmLambda)[1]
    np.testing.assert_allclose(m.prior_kl() + 0.5 * logdet, cl.prior_kl(), atol=5e-2)


@pytest.mark.parametrize("name", CASES)
def test_eval_logdet(models_golden, name):
    c = models_golden[name]
    m, cl = build(c)
    assert float(c["logdet_value"]) == 0.0
    _, KmmLambda = cl._Kmm_KmmLambda()
    grad = om.eval_logdet_grad(KmmLambda, m.conjugate_gradient)
    np.testing.assert_allclose(grad, c["logdet_grad"], rtol=1e-8, atol=1e-9)
    # reference contract (cg_test.py:72-77): gradient ~ d logdet / dA = A^-T
    np.testing.assert_allclose(grad, np.linalg.inv(KmmLambda).T, rtol=2e-2, atol=2e-2)


def test_shape_validation():
    k = g.SquaredExponential()
    with pytest.raises(ValueError):
        om.ClusterGP(k, g.Gaussian(0.1), np.zeros((4, 2)), pseudo_u=np.zeros((3, 1)))
    with pytest.raises(ValueError):
        om.ClusterGP(k, g.Gaussian(0.1), np.zeros((4, 2)), cluster_counts=np.zeros((4,)))
