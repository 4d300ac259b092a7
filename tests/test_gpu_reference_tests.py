"""The reference's own tests of this path, restated against cggp_b200 with the same shapes, thresholds and tolerances:
  cggp/cg_test.py:12-46   test_cg                    CG solution and its hyper-parameter gradients vs a direct solve
  cggp/cg_test.py:49-77   test_log_determinant_grad  eval_logdet: value 0, gradient = gradient of the true log det
(the third one, cggp/rff_test.py:9-29, is tests/test_oracle_rff.py + test_kernel_matrix_vs_reference_rff_estimates).
The reference draws un-seeded inputs; three seeds here."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def setup(seed, dimension=2, num_inputs=100):
    import cggp_b200 as cb

    rng = np.random.default_rng(seed)
    inputs = torch.as_tensor(rng.standard_normal((num_inputs, dimension))).cuda()
    lengthscales = torch.tensor(rng.random(dimension) ** 2 + 0.5, device="cuda", requires_grad=True)   # cg_test.py:22
    variance = torch.tensor(1.3, dtype=torch.float64, device="cuda", requires_grad=True)                # :23
    kernel = cb.SquaredExponential(variance=variance, lengthscales=lengthscales)
    return cb, rng, inputs, kernel, (variance, lengthscales)


def system(cb, kernel, inputs, noise_variance=0.1 ** 2):
    matrix = kernel(inputs)
    return cb.add_diagonal(matrix, noise_variance * torch.ones(matrix.shape[0], dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_cg(seed, num_systems=5, max_error=1e-12):
    cb, rng, inputs, kernel, params = setup(seed)
    rhs = torch.as_tensor(rng.standard_normal((inputs.shape[0], num_systems))).cuda()
    inv_solution = torch.linalg.solve(system(cb, kernel, inputs), rhs)
    inv_grad = torch.autograd.grad(inv_solution.sum(), params)
    cg = cb.ConjugateGradient(max_error)
    # (cg_test.py:37 hands the wrapper the TRANSPOSED right-hand sides and transposes the result back, which does not
    # fit the wrapper's documented [n, m] layout, conjugate_gradient.py:180-212; the documented layout is used here)
    cg_solution = cg(system(cb, kernel, inputs), rhs)                                                   # :36-37
    cg_grad = torch.autograd.grad(cg_solution.sum(), params)
    np.testing.assert_allclose(cg_solution.detach().cpu(), inv_solution.detach().cpu(), rtol=1e-3, atol=1e-4)  # :43
    for g1, g2 in zip(cg_grad, inv_grad):
        np.testing.assert_allclose(g1.cpu(), g2.cpu(), rtol=1e-3, atol=1e-3)                            # :45-46


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_log_determinant_grad(seed, max_error=1e-12):
    cb, rng, inputs, kernel, params = setup(seed)
    logdet = torch.logdet(system(cb, kernel, inputs))
    logdet_grad = torch.autograd.grad(logdet, params)
    logdet_cg = cb.eval_logdet(system(cb, kernel, inputs), cb.ConjugateGradient(max_error))             # :69
    logdet_grad_cg = torch.autograd.grad(logdet_cg, params)
    np.testing.assert_allclose(0, float(logdet_cg.detach()), rtol=1e-3, atol=1e-4)                             # :73
    for g1, g2 in zip(logdet_grad, logdet_grad_cg):
        np.testing.assert_allclose(g1.cpu(), g2.cpu(), rtol=1e-3, atol=1e-3)                            # :75-76
