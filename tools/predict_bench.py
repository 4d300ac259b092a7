"""BASELINE configs[3]: SGPR vs CDGP `predict_f` mean / variance on synthetic geospatial-shaped data
(N = 8M, D = 2, M = 16384, float64, 8 B200) - `python bench.py --workload c4 --mode predict` (under torchrun for N > 1).

Reference call sites: the SGPR comparison model is `gpflow.models.SGPR` built by `cggp/cli_utils.py:444-446`; the CDGP
side is `CGGP.predict_f`, `cggp/models.py:324-354`; pseudo-targets / counts come from the nearest-centre assignment of
`cggp/optimize.py:41-78`; prediction runs in batches like `cggp/cli_utils.py:426-436`.

What runs where (one process per GPU, training rows AND test points sharded over the ranks):
  CDGP   assignment (cggp_nearest_center + cggp_cluster_stats, counts / sums all-reduced) -> A = Kuu + s2 / counts
         (dense, replicated) -> per test batch ONE cggp_predict_f call (Kmn, the B-RHS CG solve on the DMMA GEMM with the
         reference's stopping rule 0.5 |r|^2 <= 1e-6, fvar / fmu reductions).
  SGPR   the system GPflow factorises, Sigma = Kuu + jitter I + Kuf Kfu / s2:  Kuf y by the fused pipelined kernel,
         Kuf Kfu [M, M] by cggp_kuf_gram (row chunks + the library's DMMA SYRK), both all-reduced.  Mean weights twice:
         (a) matrix-free preconditioned CG on the north-star operator (Nystrom preconditioner, fused tail), and
         (b) through the Cholesky factor of the dense Sigma, GPflow's own algorithm (library factorisation).
         Variance (one right-hand side per test point) through the Cholesky factors of Kuu and Sigma, as GPflow does.
The reference cannot run this size: it needs Kuf [M, N] materialised (1 TB).
"""
from __future__ import annotations

import json
import math
import os
import time

NOISE = 0.1


def _grid_inducing(M, jitter_seed=7):
    """Min-separated (cover-tree-like) inducing points: a jittered g x g grid with unit spacing, g = floor(sqrt(M))."""
    import torch

    g = int(math.isqrt(M))
    gen = torch.Generator().manual_seed(jitter_seed)
    ax = torch.arange(g, dtype=torch.float64) + 0.5
    Z = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), -1).reshape(-1, 2)
    Z = Z + (torch.rand(Z.shape, dtype=torch.float64, generator=gen) - 0.5) * 0.4
    return Z, float(g)


def _truth(X):
    import torch

    return (torch.sin(0.35 * X[:, :1]) + torch.cos(0.21 * X[:, 1:2]) * torch.sin(0.05 * X[:, :1] * X[:, 1:2] / 8.0))


def predict_compare(cb, device, rank, world, N, M, T, batch=4096, threshold=1e-6, seed=1234, timings=None,
                    matrix_free_mean=True):
    """Builds the sharded problem, runs both models on this rank's slice of the T test points.  Returns a dict with the
    predictions (device tensors), the inputs needed to re-run the models elsewhere (tests) and the timings."""
    import torch

    from cggp_b200 import _lib, selection
    from cggp_b200.sharding import shard_rows

    ctx = _lib.context(device)
    tm = timings if timings is not None else {}

    def tick():
        torch.cuda.synchronize()
        return time.perf_counter()

    Zh, side = _grid_inducing(M)
    M = Zh.shape[0]
    r_lo, r_hi = shard_rows(N, rank, world)
    g = torch.Generator(device=device).manual_seed(seed + rank)
    X = torch.rand(r_hi - r_lo, 2, dtype=torch.float64, device=device, generator=g) * side
    y = _truth(X) + math.sqrt(NOISE) * torch.randn(X.shape[0], 1, dtype=torch.float64, device=device, generator=g)
    t_lo, t_hi = shard_rows(T, rank, world)
    gt = torch.Generator(device=device).manual_seed(seed + 1000 + rank)
    Xs = torch.rand(t_hi - t_lo, 2, dtype=torch.float64, device=device, generator=gt) * side
    Z = Zh.to(device)
    kernel = cb.Matern32(variance=1.0, lengthscales=[1.0, 1.0])  # the reference's default kernel, cli_utils.py:363-368

    # ------------------------------------------------------------------ CDGP (cggp/models.py:279-354)
    t0 = tick()
    idx, _ = selection._nearest(X, Z, "sqeuclidean", None)  # optimize.py:50-51
    counts, sums = selection.cluster_stats(idx, y, M)
    if world > 1:
        ctx.allreduce_sum_(counts)
        ctx.allreduce_sum_(sums)
    empty = counts == 0
    u = torch.where(empty, torch.zeros_like(sums), sums / torch.where(empty, torch.ones_like(counts), counts))
    counts = torch.where(empty, torch.ones_like(counts), counts)  # optimize.py:70-73
    t1 = tick()
    tm["cdgp_assignment_s"] = t1 - t0
    cdgp = cb.cdgp_class(kernel, cb.Gaussian(NOISE), Z, error_threshold=threshold, cluster_counts=counts[:, None],
                         pseudo_u=u[:, None])
    mu_c, var_c, steps = [], [], []
    with torch.no_grad():
        for s in range(0, Xs.shape[0], batch):
            m_, v_ = cdgp.predict_f(Xs[s:s + batch])
            mu_c.append(m_)
            var_c.append(v_)
            steps.append(getattr(cdgp, "last_predict_steps", -1))
    t2 = tick()
    tm["cdgp_predict_s"] = t2 - t1
    tm["cdgp_cg_steps_per_batch"] = steps
    mu_c = torch.cat(mu_c) if mu_c else torch.empty((0, 1), dtype=torch.float64, device=device)
    var_c = torch.cat(var_c) if var_c else torch.empty((0, 1), dtype=torch.float64, device=device)

    # ------------------------------------------------------------------ SGPR (cli_utils.py:444-446 -> GPflow SGPR)
    t3 = tick()
    op = cb.SGPROperator(kernel, X, Z, NOISE)
    rhs = op.kuf_times(y) / NOISE  # [M, 1], all-reduced
    t4 = tick()
    tm["sgpr_operator_and_rhs_s"] = t4 - t3
    G = op.gram()  # Kuf Kfu [M, M], all-reduced
    t5 = tick()
    tm["sgpr_gram_s"] = t5 - t4
    tm["sgpr_gram_tflops_per_gpu"] = 2.0 * X.shape[0] * M * M / max(t5 - t4, 1e-9) / 1e12  # full N M^2 equivalent
    Sigma = op.Kuu + G / NOISE
    del G
    L = torch.linalg.cholesky(Sigma)   # GPflow's own algorithm factorises this system (L B L^T)
    Lk = torch.linalg.cholesky(op.Kuu)
    c_chol = torch.cholesky_solve(rhs, L)
    t6 = tick()
    tm["sgpr_factorise_s"] = t6 - t5
    mf = None
    if matrix_free_mean:
        # the north-star path for the same weights: matrix-free CG on Sigma, Nystrom preconditioner, fused tail
        pc = op.nystrom_preconditioner(num_rows=4 * M)
        t7 = tick()
        sol, (st, _) = cb.conjugate_gradient(op, rhs.t().contiguous(), None, threshold, pc, 400, 401)
        t8 = tick()
        c_cg = sol.t()
        tm["sgpr_nystrom_setup_s"] = t7 - t6
        tm["sgpr_matrix_free_solve_s"] = t8 - t7
        tm["sgpr_matrix_free_iterations"] = int(st)
        resid = Sigma @ c_cg - rhs
        mf = {"iterations": int(st), "half_rr_final": float(0.5 * (resid * resid).sum()),
              "rel_residual": float(resid.norm() / rhs.norm()),
              "max_abs_weight_diff_vs_cholesky": float((c_cg - c_chol).abs().max()),
              "it_per_s": int(st) / max(t8 - t7, 1e-9)}
        del pc
    del Sigma
    t9 = tick()
    mu_s, var_s = [], []
    for s in range(0, Xs.shape[0], batch):
        Kus = kernel.K(Z, Xs[s:s + batch])  # [M, B]
        mu_s.append(Kus.t() @ c_chol)
        t1_ = torch.linalg.solve_triangular(Lk, Kus, upper=False)
        t2_ = torch.linalg.solve_triangular(L, Kus, upper=False)
        var_s.append((kernel.K_diag(Xs[s:s + batch]) - (t1_ * t1_).sum(0) + (t2_ * t2_).sum(0))[:, None])
    t10 = tick()
    tm["sgpr_predict_s"] = t10 - t9
    mu_s = torch.cat(mu_s) if mu_s else torch.empty((0, 1), dtype=torch.float64, device=device)
    var_s = torch.cat(var_s) if var_s else torch.empty((0, 1), dtype=torch.float64, device=device)
    mu_mf = None
    if matrix_free_mean:
        mu_mf = torch.cat([kernel.K(Z, Xs[s:s + batch]).t() @ c_cg for s in range(0, Xs.shape[0], batch)]) \
            if Xs.shape[0] else mu_s.clone()
    return {"X": X, "y": y, "Z": Z, "Xs": Xs, "u": u, "counts": counts, "mu_cdgp": mu_c, "var_cdgp": var_c,
            "mu_sgpr": mu_s, "var_sgpr": var_s, "mu_sgpr_matrix_free": mu_mf, "matrix_free": mf, "truth": _truth(Xs),
            "timings": tm, "M": M, "side": side}


def run_predict(args):
    import torch
    import torch.distributed as dist

    import cggp_b200 as cb
    from bench import WORKLOADS, ClockSampler
    from cggp_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    ctx = _lib.context(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        ctx.init_comm()
    N, M, D, kern, desc = WORKLOADS[args.workload]
    if D != 2:
        raise SystemExit("--mode predict is defined for the geospatial-shaped workload (c4, D = 2)")
    if args.predict_rows:
        N = int(args.predict_rows)
    if args.predict_inducing:
        M = int(args.predict_inducing)
    T = int(args.predict_points)
    launches0 = ctx.launches
    tm = {}
    with ClockSampler(local_rank) as clk:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        res = predict_compare(cb, device, rank, world, N, M, T, batch=args.predict_batch, timings=tm)
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
    d_mu = (res["mu_sgpr"] - res["mu_cdgp"]).abs()
    d_var = (res["var_sgpr"] - res["var_cdgp"]).abs()
    d_mf = (res["mu_sgpr"] - res["mu_sgpr_matrix_free"]).abs()
    err_s = (res["mu_sgpr"] - res["truth"]) ** 2
    err_c = (res["mu_cdgp"] - res["truth"]) ** 2
    stats = torch.stack([d_mu.max(), d_var.max(), d_mf.max(), res["var_sgpr"].min(), res["var_cdgp"].min(),
                         torch.tensor(wall, dtype=torch.float64, device=device)])
    sums = torch.stack([err_s.sum(), err_c.sum(), torch.tensor(float(d_mu.numel()), dtype=torch.float64, device=device),
                        d_mu.sum(), d_var.sum()])
    mins = torch.stack([res["var_sgpr"].min(), res["var_cdgp"].min()])
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    if rank == 0:
        n_t = max(float(sums[2]), 1.0)
        line = {
            "metric": "SGPR vs CDGP predict_f mean/var (BASELINE configs[3]); seconds for the whole comparison",
            "value": float(stats[5]), "unit": "s", "higher_is_better": False, "n_gpus": world, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.workload} predict: N={N} rows sharded over {world} GPU(s), D=2, M={res['M']} "
                                   f"(jittered {int(res['side'])} x {int(res['side'])} grid, unit spacing), matern32, "
                                   f"{T} held-out test points sharded over the ranks, batches of {args.predict_batch}",
                       "cdgp": "cggp/models.py:324-354 via cggp_predict_f, error_threshold 1e-6 (cli_utils.py:439)",
                       "sgpr": "GPflow SGPR system Kuu + jitter I + Kuf Kfu / s2: Gram by cggp_kuf_gram (own DMMA SYRK), "
                               "Cholesky as in GPflow; mean weights also by matrix-free preconditioned CG"},
            "max_abs_dmean_sgpr_vs_cdgp": float(stats[0]), "max_abs_dvar_sgpr_vs_cdgp": float(stats[1]),
            "mean_abs_dmean": float(sums[3]) / n_t, "mean_abs_dvar": float(sums[4]) / n_t,
            "max_abs_dmean_matrix_free_vs_cholesky": float(stats[2]),
            "rmse_vs_noise_free_truth": {"sgpr": math.sqrt(float(sums[0]) / n_t), "cdgp": math.sqrt(float(sums[1]) / n_t)},
            "min_variance": {"sgpr": float(mins[0]), "cdgp": float(mins[1])},
            "matrix_free_mean_solve": res["matrix_free"],
            "timings_rank0_s": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in tm.items()},
            "gpu_launches": int(ctx.launches - launches0),
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
