"""Parity at BASELINE.json's FULL sizes (c3: N=2M, M=4096, D=11, Matern-5/2; c2: N=434 874, M=2048, D=3, SE) through
size-independent properties of the operator Sigma = Kuu + Kuf Kfu / s2 - the oracle cannot run at these sizes in
seconds, the properties can: symmetry, linearity, positive semi-definiteness, additivity over row shards (the
multi-GPU decomposition), agreement of the fused kernel with the independent two-sweep kernels on a row sample,
bitwise reproducibility, and monotone decrease of the CG energy norm."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CONFIGS = {"c3": (2_000_000, 4096, 11, "matern52"), "c2": (434_874, 2048, 3, "se")}


@pytest.fixture(scope="module")
def cb():
    import cggp_b200

    return cggp_b200


@pytest.fixture(scope="module", params=["c3", "c2"])
def problem(request, cb):
    N, M, D, kern = CONFIGS[request.param]
    gen = torch.Generator(device="cuda").manual_seed(123)
    X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=gen)
    Z = X[torch.randperm(N, device="cuda", generator=gen)[:M]].clone()
    k = cb.kernels.KERNELS[kern](variance=1.0, lengthscales=[1.0] * D)
    op = cb.SGPROperator(k, X, Z, 0.1)
    U = torch.randn(2, M, dtype=torch.float64, device="cuda", generator=gen)
    return request.param, op, X, Z, k, U


def test_symmetry_linearity_psd(cb, problem):
    name, op, X, Z, k, U = problem
    W = op.kuf_kfu_matmul(U)  # both right-hand sides in one launch (NB = 2 plan)
    w0 = op.kuf_kfu_matmul(U[:1])  # NB = 1 plan
    np.testing.assert_allclose(w0.cpu().numpy(), W[:1].cpu().numpy(), rtol=1e-12)
    # symmetry: <u1, S u0> == <u0, S u1>
    a, b = float((U[1] * W[0]).sum()), float((U[0] * W[1]).sum())
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b), float(W.abs().max()))
    # linearity
    comb = op.kuf_kfu_matmul(2.5 * U[:1] - 0.5 * U[1:])
    np.testing.assert_allclose(comb.cpu().numpy(), (2.5 * W[:1] - 0.5 * W[1:]).cpu().numpy(), rtol=1e-10,
                               atol=1e-12 * float(W.abs().max()))
    # positive semi-definite quadratic form: u (Kuf Kfu) u^T = |Kfu u|^2
    assert float((U[0] * W[0]).sum()) > 0 and float((U[1] * W[1]).sum()) > 0
    # bitwise reproducible
    assert torch.equal(W, op.kuf_kfu_matmul(U))


def test_additivity_over_row_shards_and_two_sweep_agreement(cb, problem):
    name, op, X, Z, k, U = problem
    N = X.shape[0]
    full = op.kuf_kfu_matmul(U[:1])
    cut = (N // 3) + 5  # ragged split: the partial products of disjoint row shards add up to the full product
    parts = [cb.SGPROperator(k, X[s:e], Z, 0.1).kuf_kfu_matmul(U[:1]) for s, e in ((0, cut), (cut, N))]
    np.testing.assert_allclose((parts[0] + parts[1]).cpu().numpy(), full.cpu().numpy(), rtol=1e-11,
                               atol=1e-12 * float(full.abs().max()))
    # fused pipelined kernel vs the independent two-sweep kernels (validated against the oracle at small sizes)
    sub = cb.SGPROperator(k, X[: 96_000 + 7], Z, 0.1)
    np.testing.assert_allclose(sub.kuf_kfu_matmul(U[:1], variant=3).cpu().numpy(),
                               sub.kuf_kfu_matmul(U[:1], variant=1).cpu().numpy(), rtol=1e-11,
                               atol=1e-12 * float(full.abs().max()))


def test_cg_energy_decreases_and_preconditioned_solve_converges(cb, problem):
    name, op, X, Z, k, U = problem
    gen = torch.Generator(device="cuda").manual_seed(5)
    y = torch.sin(X.sum(-1, keepdim=True)) + 0.3 * torch.randn(X.shape[0], 1, dtype=torch.float64, device="cuda",
                                                                 generator=gen)
    rhs = (op.kuf_times(y) / 0.1).t().contiguous()
    # CG minimises the energy 0.5 x S x - b x monotonically over the iterations (exact-arithmetic property that
    # survives rounding for the first iterations of any SPD system)
    energies = []
    for its in (1, 2, 4, 8):
        x, (steps, _) = cb.conjugate_gradient(op, rhs, None, 0.0, None, its, its + 1)
        assert int(steps) == its
        energies.append(float(0.5 * (x * op.matmul(x)).sum() - (x * rhs).sum()))
    assert all(e1 < e0 for e0, e1 in zip(energies, energies[1:])), energies
    # full solve with the Nystrom preconditioner to the reference's absolute threshold (cli_utils.py:439)
    pc = op.nystrom_preconditioner()
    sol, (steps, err, hist) = cb.conjugate_gradient(op, rhs, None, 1e-6, pc, 400, 401, return_history=True)
    assert int(steps) < 400 and float(hist[-1].max()) <= 1e-6
    resid = op.matmul(sol) - rhs
    assert float(0.5 * (resid * resid).sum()) < 1e-4  # true residual agrees with the recursive one
