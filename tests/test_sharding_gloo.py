"""N > 1 host logic on CPU (gloo, world_size 2): the row sharding partitions the data, and the sharded data term
sum_r Kuf_r (Kfu_r v) combined by one all-reduce equals the unsharded product.  The per-rank compute
here is the oracle (no GPU in this container); the CUDA ranks are exercised by tests/test_multi_gpu.py on the box."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, M, D, out):
    sys.path.insert(0, ROOT)
    from cggp_b200.sharding import shard_rows
    from oracle import cg as ocg
    from oracle import gpflow_restated as g
    from oracle import models as om

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # every rank builds the same global problem and takes its rows
    X, Z = rng.standard_normal((N, D)), rng.standard_normal((M, D))
    y = rng.standard_normal((N, 1))
    k = g.Matern52(variance=1.1, lengthscales=np.full(D, 1.3))
    s, e = shard_rows(N, rank, world)
    Kuu = g.Kuu(Z, k, 1e-6)

    def matmul(V):  # Sigma = Kuu + Kuf Kfu / s2 with the data term all-reduced (the one collective of the path)
        # the fused tail kernel's form (csrc/cg.cu cg_tail_kernel): every rank adds ITS columns [r M / W, (r + 1) M / W)
        # of the replicated term V Kuu, divided by the scale 1 / s2, to its partial product; the one sum over ranks then
        # delivers V Sigma / scale
        scale = 1.0 / 0.1
        part = om.kuf_kfu_matmul(k, X[s:e], Z, V)
        lo, hi = M * rank // world, M * (rank + 1) // world
        part[:, lo:hi] += (V @ Kuu[:, lo:hi]) / scale
        part = torch.from_numpy(part)
        dist.all_reduce(part)
        return scale * part.numpy()

    rhs_part = torch.from_numpy((k.K(Z, X[s:e]) @ y[s:e] / 0.1).T.copy())
    dist.all_reduce(rhs_part)
    rhs = rhs_part.numpy()
    hist = []
    sol, (steps, _) = ocg.conjugate_gradient(matmul, rhs, np.zeros_like(rhs), 0.0, None, 8, 100, history=hist)
    if rank == 0:
        full = om.sgpr_operator(k, X, Z, 0.1)
        rhs_full = (k.K(Z, X) @ y / 0.1).T
        hist_full = []
        sol_full, _ = ocg.conjugate_gradient(full, rhs_full, np.zeros_like(rhs_full), 0.0, None, 8, 100,
                                             history=hist_full)
        dev = np.abs(np.array(hist) / np.array(hist_full) - 1).max(axis=1)  # per iteration
        np.save(out, np.array([dev[:4].max(), dev.max(),
                               np.abs(sol - sol_full).max() / np.abs(sol_full).max(), float(e - s)]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows_partitions():
    sys.path.insert(0, ROOT)
    from cggp_b200.sharding import shard_rows

    for n in (0, 1, 7, 2_000_000, 434_874):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def test_sharded_operator_allreduce_matches_unsharded(tmp_path):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), 1501, 48, 3, out), nprocs=2, join=True)
    dev_early, dev_all, dev_sol, n0 = np.load(out)
    assert n0 == 751  # first rank takes the extra row
    # sharding only changes the summation order: 1e-9 while CG has not amplified the rounding (the system has
    # cond ~ 1e9), loose afterwards
    assert dev_early < 1e-9 and dev_all < 1e-5 and dev_sol < 1e-5


def test_shard_rows_weighted_partitions_and_follows_the_speeds():
    from cggp_b200.sharding import shard_rows, shard_rows_weighted

    for n in (0, 5, 1000, 250_007, 2_000_000):
        for speeds in ([1.0], [1.0, 1.0], [1.0, 0.98, 1.01, 0.99, 1.0, 1.02, 0.97, 1.0], [3.0, 1.0, 1.0]):
            world = len(speeds)
            spans = [shard_rows_weighted(n, r, speeds) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            assert all(a <= b for a, b in spans)
            if n >= 100_000:
                total = sum(speeds)
                for (a, b), s in zip(spans, speeds):
                    assert abs((b - a) / n - s / total) < 1e-3
    # equal speeds: the even split up to the block granularity
    for r in range(4):
        a, b = shard_rows_weighted(2_000_000, r, [1.0] * 4)
        e = shard_rows(2_000_000, r, 4)
        assert abs(a - e[0]) <= 24 and abs(b - e[1]) <= 24
    with pytest.raises(ValueError):
        shard_rows_weighted(10, 0, [1.0, 0.0])
