"""`cggp_b200.from_gpflow`: GPflow kernels / likelihoods / inducing variables (recognised by duck typing; here the
stand-in classes of the golden shim and a minimal Parameter with `.numpy()`) -> the objects of this package, and the
models accept them directly as the reference's do (cggp/models.py:279-291)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "_shim"))

torch = pytest.importorskip("torch")


class FakeParameter:
    """What gpflow.Parameter offers to a reader: `.numpy()` (no __dlpack__, no torch)."""

    def __init__(self, value):
        self._v = np.asarray(value, dtype=np.float64)

    def numpy(self):
        return self._v


def shim():
    import gpflow  # the shim under tests/golden/_shim

    return gpflow


def test_kernels_and_likelihood_convert_by_duck_typing():
    import cggp_b200 as cb

    gp = shim()
    for name, cls in (("SquaredExponential", cb.SquaredExponential), ("Matern12", cb.Matern12),
                      ("Matern32", cb.Matern32), ("Matern52", cb.Matern52)):
        gk = getattr(gp.kernels, name)(variance=1.7, lengthscales=np.array([0.5, 2.0, 1.25]))
        k = cb.from_gpflow(gk)
        assert type(k) is cls and k.variance == 1.7
        np.testing.assert_array_equal(k.lengthscales.numpy(), [0.5, 2.0, 1.25])
        assert cb.from_gpflow(k) is k  # objects of this package pass through

    class Matern32:  # a real GPflow kernel exposes Parameters, not arrays
        variance = FakeParameter(0.3)
        lengthscales = FakeParameter([1.5])
        active_dims = slice(None)

    k = cb.from_gpflow(Matern32())
    assert type(k) is cb.Matern32 and k.variance == 0.3 and k.lengthscales.tolist() == [1.5]
    lik = cb.from_gpflow(gp.likelihoods.Gaussian(0.25))
    assert isinstance(lik, cb.Gaussian) and lik.variance == 0.25

    class Gaussian:
        variance = FakeParameter(0.1)

    assert cb.from_gpflow(Gaussian()).variance == 0.1

    class Periodic:
        variance, lengthscales = 1.0, 1.0

    with pytest.raises(TypeError):
        cb.from_gpflow(Periodic())


@pytest.mark.gpu
def test_models_accept_gpflow_objects():
    import cggp_b200 as cb

    gp = shim()
    rng = np.random.default_rng(0)
    M, D = 40, 2
    Z = rng.standard_normal((M, D))
    Xs = rng.standard_normal((17, D))
    u = rng.standard_normal((M, 1))
    counts = rng.integers(1, 9, (M, 1)).astype(np.float64)
    gk = gp.kernels.Matern52(variance=1.3, lengthscales=np.array([0.9, 1.4]))
    giv = gp.models.util.inducingpoint_wrapper(Z)
    glik = gp.likelihoods.Gaussian(0.2)
    cg = cb.ConjugateGradient(1e-14)
    m_gp = cb.CGGP(gk, glik, giv, cg, num_probes=None, pseudo_u=u, cluster_counts=counts)
    m_nat = cb.CGGP(cb.Matern52(1.3, [0.9, 1.4]), cb.Gaussian(0.2), torch.as_tensor(Z).cuda(), cg, num_probes=None,
                    pseudo_u=u, cluster_counts=counts)
    assert isinstance(m_gp.kernel, cb.Matern52) and isinstance(m_gp.inducing_variable, cb.InducingPoints)
    mu1, v1 = m_gp.predict_f(Xs)
    mu2, v2 = m_nat.predict_f(Xs)
    assert torch.equal(mu1, mu2) and torch.equal(v1, v2)
    assert torch.equal(m_gp.prior_kl(), m_nat.prior_kl())
    # the SGPR factory of cli_utils.py:444-446 with a GPflow kernel / likelihood / inducing variable
    X, Y = rng.standard_normal((300, D)), rng.standard_normal((300, 1))
    s1 = cb.sgpr_class((X, Y), gk, glik, giv, conjugate_gradient=cb.ConjugateGradient(1e-12, max_iterations=300))
    s2 = cb.sgpr_class((X, Y), cb.Matern52(1.3, [0.9, 1.4]), cb.Gaussian(0.2), Z,
                       conjugate_gradient=cb.ConjugateGradient(1e-12, max_iterations=300))
    assert torch.equal(s1.predict_f(Xs)[0], s2.predict_f(Xs)[0])
