"""oracle.cg vs golden vectors produced by the reference's own cggp/conjugate_gradient.py (tests/golden)."""
import numpy as np
import pytest

from oracle import cg as ocg

CASES = ["cgtest_se", "matern32_thr1e-6", "reset_cycle7", "maxit_cap", "zero_rhs_row", "float32"]


@pytest.mark.parametrize("name", CASES)
def test_cg_matches_reference_run(cg_golden, name):
    c = cg_golden[name]
    max_it = None if int(c["max_it"]) < 0 else int(c["max_it"])
    hist = []
    sol, (steps, err) = ocg.conjugate_gradient(c["A"], c["rhs"], c["x0"], float(c["thr"]), None, max_it,
                                               int(c["cycle"]), history=hist)
    # same arithmetic, same op order -> bit-identical on the same BLAS; allow a few ulp for BLAS variation
    tol = 1e-4 if c["A"].dtype == np.float32 else 1e-10
    assert int(steps) == int(c["steps"])
    assert sol.dtype == c["solution"].dtype
    np.testing.assert_allclose(sol, c["solution"], rtol=tol, atol=tol)
    np.testing.assert_allclose(err, c["error"], rtol=1e-6, atol=1e-30)
    np.testing.assert_allclose(np.array(hist), c["history"], rtol=1e-6, atol=1e-30)


@pytest.mark.parametrize("name", ["cgtest_se", "matern32_thr1e-6", "reset_cycle7"])
def test_cg_backward_matches_reference_closure(cg_golden, name):
    c = cg_golden[name]
    max_it = None if int(c["max_it"]) < 0 else int(c["max_it"])
    dA, db = ocg.grad_conjugate_gradient(c["A"], c["solution"], c["dx"], float(c["thr"]), None, max_it,
                                         int(c["cycle"]))
    np.testing.assert_allclose(db, c["db"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(dA, c["dA"], rtol=1e-9, atol=1e-9)


def test_adapter_layout(cg_golden):
    c = cg_golden["adapter"]
    cg = ocg.ConjugateGradient(float(c["thr"]))
    sol = cg(c["A"], c["rhs"])
    assert sol.shape == c["rhs"].shape
    np.testing.assert_allclose(sol, c["solution"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(cg.last_history, c["history"], rtol=1e-6, atol=1e-30)


def test_reference_test_contract_cg_vs_solve(cg_golden):
    """cggp/cg_test.py:34-43: CG(1e-12) ~ direct solve at rtol 1e-3 / atol 1e-4."""
    c = cg_golden["cgtest_se"]
    direct = np.linalg.solve(c["A"], c["rhs"].T).T
    np.testing.assert_allclose(c["solution"], direct, rtol=1e-3, atol=1e-4)


def test_callable_matrix_equals_dense(cg_golden):
    c = cg_golden["matern32_thr1e-6"]
    A = c["A"]
    s1, st1 = ocg.conjugate_gradient(A, c["rhs"], c["x0"], 1e-6)
    s2, st2 = ocg.conjugate_gradient(lambda V: V @ A, c["rhs"], c["x0"], 1e-6)
    assert int(st1[0]) == int(st2[0])
    np.testing.assert_array_equal(s1, s2)


def test_block_preconditioner():
    rng = np.random.default_rng(0)
    n, bs = 96, 12
    X = np.sort(rng.uniform(0, 10, size=(n, 1)), axis=0)
    A = np.exp(-0.5 * (X - X.T) ** 2) + 1e-2 * np.eye(n)
    rhs = rng.standard_normal((2, n))
    blocks = np.arange(n).reshape(n // bs, bs)
    sol, (steps, _) = ocg.conjugate_gradient(A, rhs, np.zeros_like(rhs), 1e-10, ocg.BlockPreconditioner(blocks), 500,
                                             1000)
    assert steps < 500
    np.testing.assert_allclose(sol @ A, rhs, atol=1e-4)
    # one block covering everything = exact inverse: converges in one step
    sol1, (s1, _) = ocg.conjugate_gradient(A, rhs, np.zeros_like(rhs), 1e-10,
                                           ocg.BlockPreconditioner(np.arange(n)[None, :]), 500, 1000)
    assert int(s1) <= 2
    np.testing.assert_allclose(sol1 @ A, rhs, atol=1e-5)
