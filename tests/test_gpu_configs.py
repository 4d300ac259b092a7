"""The five BASELINE.json configs as parity cases (reduced replicas where the oracle cannot run the full size in
seconds; the full-size properties are in tests/test_gpu_fullsize.py):
  configs[0] CDGP CG solve, 2-D regression N=10k, M=500 uniform inducing points, SE, float64 - FULL size, end to end
  configs[1] CDGP ELBO + CG solve, 3droad-shaped (D=3), SE - replica N=40k, M=256 vs oracle + exact Cholesky twin
  configs[2] CG solve on houseelectric-shaped data (D=11, Matern-5/2) - replica N=30k, M=256 (headline kernel path)
  configs[3] SGPR vs CDGP predict_f, geospatial-shaped (D=2) - replica N=20k, M=256
  configs[4] float32 solve, D=90, SE - replica N=6k, M=192, tolerance 1e-4"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cg as ocg  # noqa: E402
from oracle import gpflow_restated as g  # noqa: E402
from oracle import models as om  # noqa: E402
from oracle import noise as nz  # noqa: E402


def dev(x, dtype=None):
    return torch.as_tensor(np.asarray(x), dtype=dtype).cuda()


def cpu(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def cb():
    import cggp_b200

    return cggp_b200


def test_config1_cdgp_full_size_end_to_end(cb):
    """X ~ U[-5,5]^2, N=10 000, Z = 500 rows of X, SE(variance 1, lengthscale 1), noise 0.1, threshold 1e-6
    (cli_utils.py:153,363-368,439): nearest-centre statistics -> CDGP predict_f / prior_kl / elbo, every piece on the
    device, against the oracle pipeline."""
    rng = np.random.default_rng(0)
    N, M = 10_000, 500
    X = rng.uniform(-5, 5, (N, 2))
    y = np.sin(X[:, :1]) * np.cos(X[:, 1:]) + np.sqrt(0.1) * rng.standard_normal((N, 1))
    Z = X[rng.choice(N, M, replace=False)].copy()
    ok, k = g.SquaredExponential(1.0, [1.0, 1.0]), cb.SquaredExponential(1.0, [1.0, 1.0])
    # assignment + cluster statistics (optimize.py:41-78 semantics)
    _, omeans, ocounts = om.oips_style_assignment(Z, X, y)
    _, means, counts = cb.selection.nearest_center_update(dev(Z), (dev(X), dev(y)))
    np.testing.assert_array_equal(cpu(counts), ocounts.astype(np.float64))
    np.testing.assert_allclose(cpu(means), omeans, rtol=1e-12, equal_nan=True)
    u = np.nan_to_num(omeans)[:, None]
    cnt = ocounts.astype(np.float64)[:, None]
    probes = np.sign(rng.standard_normal((M, 5)))
    m = cb.cdgp_class(k, cb.Gaussian(0.1), dev(Z), error_threshold=1e-6, cluster_counts=dev(cnt), pseudo_u=dev(u),
                      num_data=N)
    m.probes = dev(probes)
    Xs = X[:1000]
    mu, var = m.predict_f(dev(Xs))
    kl = float(m.prior_kl())
    elbo = float(m.elbo((dev(X[:1000]), dev(y[:1000]))))

    def oracle(cgobj):
        mo = om.CGGP(ok, g.Gaussian(0.1), Z, cgobj, cluster_counts=cnt, pseudo_u=u, num_data=N)
        mo.probes = probes
        omu, ovar = mo.predict_f(Xs)
        return omu, ovar, mo.prior_kl(), mo.elbo((X[:1000], y[:1000]))

    ref = oracle(ocg.ConjugateGradient(1e-6))
    alts = [oracle(nz.PermutedCG(s, 1e-6)) for s in (0, 1)]
    for got, i, what in ((cpu(mu), 0, "mean"), (cpu(var), 1, "var"), (kl, 2, "kl"), (elbo, 3, "elbo")):
        nz.assert_close_with_noise(got, ref[i], [a[i] for a in alts], 1e-8, what)
    # the CG answer is also close to the exact (Cholesky) twin of the same model
    cl = cb.ClusterGP(k, cb.Gaussian(0.1), dev(Z), cluster_counts=dev(cnt), pseudo_u=dev(u), num_data=N)
    cmu, cvar = cl.predict_f(dev(Xs))
    assert float((mu - cmu).abs().max()) < 1e-2 and float((var - cvar).abs().max()) < 1e-2


def test_config2_cdgp_elbo_replica(cb):
    rng = np.random.default_rng(1)
    N, M, D = 40_000, 256, 3
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(-1, keepdims=True)) + np.sqrt(0.1) * rng.standard_normal((N, 1))
    # min-separation (cover-tree-like) inducing points: greedy epsilon-net on a subsample
    cand = X[rng.choice(N, 4000, replace=False)]
    Zl = [cand[0]]
    for c in cand[1:]:
        if len(Zl) == M:
            break
        if np.min(np.sum((np.array(Zl) - c) ** 2, -1)) > 0.35 ** 2:
            Zl.append(c)
    Z = np.array(Zl)
    M = Z.shape[0]
    ok, k = g.SquaredExponential(1.0, np.ones(D)), cb.SquaredExponential(1.0, np.ones(D))
    _, omeans, ocounts = om.oips_style_assignment(Z, X, y)
    u, cnt = np.nan_to_num(omeans)[:, None], ocounts.astype(np.float64)[:, None]
    thr = 1e-10
    m = cb.CGGP(k, cb.Gaussian(0.1), dev(Z), cb.ConjugateGradient(thr), num_probes=None, cluster_counts=dev(cnt),
                pseudo_u=dev(u), num_data=N)
    batch = (X[:2000], y[:2000])
    elbo = float(m.elbo((dev(batch[0]), dev(batch[1]))))

    def oracle(cgobj):
        mo = om.CGGP(ok, g.Gaussian(0.1), Z, cgobj, num_probes=None, cluster_counts=cnt, pseudo_u=u, num_data=N)
        return mo.elbo(batch)

    ref = oracle(ocg.ConjugateGradient(thr))
    nz.assert_close_with_noise(elbo, ref, [oracle(nz.PermutedCG(s, thr)) for s in (0, 1)], 1e-8, "elbo")
    # exact twin: CGGP's KL lacks log det (Kmm + Lambda) (eval_logdet's forward value is 0, models.py:46,319)
    cl = cb.ClusterGP(k, cb.Gaussian(0.1), dev(Z), cluster_counts=dev(cnt), pseudo_u=dev(u), num_data=N)
    KmmL = cb.add_diagonal(cb.Kuu(dev(Z), k), cl.diag_variance[:, 0])
    logdet = float(torch.linalg.slogdet(KmmL)[1])
    # (trace over M CG solves, each stopped at 0.5|r|^2 <= 1e-10: agreement to ~1e-5)
    np.testing.assert_allclose(float(m.prior_kl()) + 0.5 * logdet, float(cl.prior_kl()), rtol=2e-4)


def test_config2_cover_tree_selected_inducing_points(cb):
    """configs[1] as worded: "M cover-tree-selected" - the inducing points, pseudo targets and counts come from the
    device cover tree (optimize.py:19-39), bit-identical to the oracle's tree, and feed the CDGP ELBO."""
    import warnings

    from oracle import covertree as oct_

    rng = np.random.default_rng(11)
    N, D = 30_000, 3
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(-1, keepdims=True)) + np.sqrt(0.1) * rng.standard_normal((N, 1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = oct_.CoverTree(None, (X, y), spatial_resolution=0.8)
    omeans, ocounts = tree.cluster_mean_and_counts
    keep = ocounts.reshape(-1) != 0
    Z, u, cnt = tree.centroids[keep], omeans[keep], ocounts[keep]
    iv, means, counts = cb.covertree_update_inducing_parameters(None, (dev(X), dev(y)), None, 0.8)
    np.testing.assert_array_equal(cpu(iv), Z)
    np.testing.assert_array_equal(cpu(means), u)
    np.testing.assert_array_equal(cpu(counts), cnt)
    ok, k = g.SquaredExponential(1.0, np.ones(D)), cb.SquaredExponential(1.0, np.ones(D))
    thr = 1e-10
    m = cb.CGGP(k, cb.Gaussian(0.1), iv, cb.ConjugateGradient(thr), num_probes=None, cluster_counts=counts,
                pseudo_u=means, num_data=N)
    batch = (X[:2000], y[:2000])
    elbo = float(m.elbo((dev(batch[0]), dev(batch[1]))))

    def oracle(cgobj):
        mo = om.CGGP(ok, g.Gaussian(0.1), Z, cgobj, num_probes=None, cluster_counts=cnt, pseudo_u=u, num_data=N)
        return mo.elbo(batch)

    ref = oracle(ocg.ConjugateGradient(thr))
    nz.assert_close_with_noise(elbo, ref, [oracle(nz.PermutedCG(s, thr)) for s in (0, 1)], 1e-8, "elbo")


def test_config3_headline_solve_replica(cb):
    """Matrix-free CG on Sigma = Kuu + Kuf Kfu / s2, D=11, Matern-5/2, through the pipelined kernel (variant 3, both
    the B=1 and B=2 plans), vs the oracle: trajectory, iteration count, Nystrom-preconditioned full solve."""
    rng = np.random.default_rng(2)
    N, M, D = 30_000, 256, 11
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(-1, keepdims=True)) + np.sqrt(0.1) * rng.standard_normal((N, 2))
    Z = X[rng.choice(N, M, replace=False)] + 0.01
    ls = np.full(D, 2.0)
    ok, k = g.Matern52(1.0, ls), cb.Matern52(1.0, ls)
    op = cb.SGPROperator(k, dev(X), dev(Z), 0.1, variant=3)
    rhs = (ok.K(Z, X) @ y / 0.1).T  # [2, M]
    np.testing.assert_allclose(cpu(op.kuf_times(dev(y)) / 0.1).T, rhs, rtol=1e-11)
    hist, alts = [], []
    ocg.conjugate_gradient(om.sgpr_operator(ok, X, Z, 0.1), rhs, np.zeros_like(rhs), 0.0, None, 15, 1000, history=hist)
    for chunk in (1000, 3333):
        h = []
        ocg.conjugate_gradient(om.sgpr_operator(ok, X, Z, 0.1, chunk=chunk), rhs, np.zeros_like(rhs), 0.0, None, 15,
                               1000, history=h)
        alts.append(np.array(h))
    for B in (1, 2):
        _, (steps, _, h) = cb.conjugate_gradient(op, dev(rhs[:B]), None, 0.0, None, 15, 1000, return_history=True)
        assert int(steps) == 15
        k_ = min(len(h), len(hist))
        dev_ = np.abs(cpu(h)[:k_] - np.array(hist)[:k_, :B]) / np.array(hist)[:k_, :B]
        floor = np.maximum.accumulate(np.max([np.abs(a[:k_, :B] - np.array(hist)[:k_, :B]) / np.array(hist)[:k_, :B]
                                              for a in alts], axis=0), axis=0)
        assert (dev_ <= np.maximum(2e-9, 50 * floor)).all(), dev_.max(axis=1)
        assert (dev_[:3] <= 2e-9).all()


def test_config4_sgpr_vs_cdgp_predict_replica(cb):
    rng = np.random.default_rng(3)
    N, M, D = 20_000, 256, 2
    X = rng.uniform(0, 16, (N, D))
    y = np.sin(X[:, :1]) + np.cos(0.5 * X[:, 1:]) + np.sqrt(0.1) * rng.standard_normal((N, 1))
    gx = np.linspace(0.5, 15.5, 16)
    Z = np.stack(np.meshgrid(gx, gx), -1).reshape(-1, 2)  # spacing = lengthscale
    Xs = rng.uniform(0, 16, (300, D))
    ok, k = g.Matern52(1.0, [1.0, 1.0]), cb.Matern52(1.0, [1.0, 1.0])
    # SGPR through the matrix-free system with the Nystrom preconditioner vs GPflow's Cholesky formulas
    ref_mean, ref_var = g.SGPR((X, y), ok, Z, noise_variance=0.1).predict_f(Xs)
    model = cb.sgpr_class((dev(X), dev(y)), k, cb.Gaussian(0.1), dev(Z))
    pc = model.operator.nystrom_preconditioner()
    model.conjugate_gradient = cb.ConjugateGradient(1e-16, preconditioner=pc, max_iterations=500)
    mean, var = model.predict_f(dev(Xs))
    np.testing.assert_allclose(cpu(mean), ref_mean, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cpu(var), ref_var, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(model.elbo()), g.SGPR((X, y), ok, Z, noise_variance=0.1).elbo(), rtol=1e-9)
    # CDGP on the same data: cluster means as pseudo-targets; close to SGPR where the clusters are small
    _, omeans, ocounts = om.oips_style_assignment(Z, X, y)
    u, cnt = np.nan_to_num(omeans)[:, None], ocounts.astype(np.float64)[:, None]
    cd = cb.cdgp_class(k, cb.Gaussian(0.1), dev(Z), error_threshold=1e-12, cluster_counts=dev(cnt), pseudo_u=dev(u))
    cmu, cvar = cd.predict_f(dev(Xs))
    mo = om.CGGP(ok, g.Gaussian(0.1), Z, ocg.ConjugateGradient(1e-12), cluster_counts=cnt, pseudo_u=u)
    omu, ovar = mo.predict_f(Xs)
    alt = [om.CGGP(ok, g.Gaussian(0.1), Z, nz.PermutedCG(s, 1e-12), cluster_counts=cnt, pseudo_u=u).predict_f(Xs)
           for s in (0, 1)]
    nz.assert_close_with_noise(cpu(cmu), omu, [a[0] for a in alt], 1e-8, "cdgp mean")
    nz.assert_close_with_noise(cpu(cvar), ovar, [a[1] for a in alt], 1e-8, "cdgp var")
    assert float(np.abs(cpu(cmu) - ref_mean).mean()) < 0.2  # two approximations of the same posterior


def test_config4_predict_bench_code_path_vs_oracle(cb):
    """The code path `bench.py --workload c4 --mode predict` times at N = 8M / M = 16384 (tools/predict_bench.py:
    assignment -> cggp_predict_f batches; Kuf y + cggp_kuf_gram -> Sigma -> mean / variance; matrix-free preconditioned
    CG for the mean weights) on the reduced replica SURVEY.md 8(d) prescribes, against the oracle's CGGP and GPflow-SGPR
    restatements on the same inputs."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools.predict_bench import predict_compare

    res = predict_compare(cb, torch.device("cuda", 0), 0, 1, N=100_000, M=1024, T=600, batch=256, threshold=1e-14)
    X, y, Z, Xs = cpu(res["X"]), cpu(res["y"]), cpu(res["Z"]), cpu(res["Xs"])
    u, cnt = cpu(res["u"])[:, None], cpu(res["counts"])[:, None]
    ok = g.Matern32(1.0, [1.0, 1.0])
    # assignment: same clusters as the oracle's (optimize.py:50-78 semantics)
    _, omeans, ocounts = om.oips_style_assignment(Z, X, y)
    np.testing.assert_array_equal(cnt[:, 0], ocounts.astype(np.float64))
    np.testing.assert_allclose(u[:, 0], np.nan_to_num(omeans), rtol=1e-10, atol=1e-12)
    # CDGP
    mo = om.CGGP(ok, g.Gaussian(0.1), Z, ocg.ConjugateGradient(1e-14), cluster_counts=cnt, pseudo_u=u)
    omu, ovar = mo.predict_f(Xs)
    alt = [om.CGGP(ok, g.Gaussian(0.1), Z, nz.PermutedCG(s, 1e-14), cluster_counts=cnt, pseudo_u=u).predict_f(Xs)
           for s in (0, 1)]
    nz.assert_close_with_noise(cpu(res["mu_cdgp"]), omu, [a[0] for a in alt], 1e-8, "cdgp mean")
    nz.assert_close_with_noise(cpu(res["var_cdgp"]), ovar, [a[1] for a in alt], 1e-8, "cdgp var")
    # SGPR (GPflow's Cholesky formulas)
    ref_mean, ref_var = g.SGPR((X, y), ok, Z, noise_variance=0.1).predict_f(Xs)
    np.testing.assert_allclose(cpu(res["mu_sgpr"]), ref_mean, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cpu(res["var_sgpr"]), ref_var, rtol=1e-5, atol=1e-6)
    # the matrix-free preconditioned CG reaches the same mean weights
    assert float((res["mu_sgpr_matrix_free"] - res["mu_sgpr"]).abs().max()) < 1e-5
    assert float((res["mu_sgpr"] - res["mu_cdgp"]).abs().mean()) < 0.2  # two approximations of one posterior


def test_config5_float32_replica(cb):
    rng = np.random.default_rng(4)
    N, M, D = 6000, 192, 90
    X = rng.standard_normal((N, D)).astype(np.float32)
    Z = X[rng.choice(N, M, replace=False)].copy()
    y = np.sin(X[:, :3].sum(-1, keepdims=True)).astype(np.float32)
    ls = np.full(D, np.sqrt(D), np.float32)  # r^2 = O(1)
    ok = g.SquaredExponential(1.0, ls, dtype=np.float32)
    k = cb.SquaredExponential(1.0, ls)
    Kd = cpu(k.K(dev(X[:500]), dev(Z)))
    assert Kd.dtype == np.float32
    np.testing.assert_allclose(Kd, ok.K(X[:500], Z), rtol=1e-4, atol=1e-6)
    op = cb.SGPROperator(k, dev(X), dev(Z), 0.1)
    V = rng.standard_normal((2, M)).astype(np.float32)
    ref = om.kuf_kfu_matmul(g.SquaredExponential(1.0, ls.astype(np.float64)), X.astype(np.float64),
                            Z.astype(np.float64), V.astype(np.float64))
    W = cpu(op.kuf_kfu_matmul(dev(V)))
    assert W.dtype == np.float32
    np.testing.assert_allclose(W, ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())
    # dense float32 CG (the reference's guard constant stays 1e-16 in float32, conjugate_gradient.py:50)
    A = (ok.K(Z) + 0.1 * np.eye(M, dtype=np.float32)).astype(np.float32)
    rhs = rng.standard_normal((3, M)).astype(np.float32)
    # threshold above the float32 rounding floor of this system, so that the iteration count is well defined
    osol, (osteps, _) = ocg.conjugate_gradient(A, rhs, np.zeros_like(rhs), 1e-4, None, None, 100)
    sol, (steps, _) = cb.conjugate_gradient(dev(A), dev(rhs), None, 1e-4, None, None, 100)
    assert sol.dtype == torch.float32 and abs(int(steps) - int(osteps)) <= 1
    np.testing.assert_allclose(cpu(sol), osol, rtol=1e-3, atol=1e-3)
