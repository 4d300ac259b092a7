"""Multi-GPU parity check (run under torchrun, one rank per GPU): the sharded matrix-free CG (partial Kuf_r Kfu_r v per
rank + one NCCL all-reduce per iteration) against the same solve on one GPU with all rows.  Rank 0 prints one line
`MGPU_CHECK ok ...` or raises.  Used by tests/test_multi_gpu.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import cggp_b200 as cb
from cggp_b200 import _lib
from cggp_b200.sharding import shard_rows


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.context(dev)
    N, M, D, its = 200_003, 1024, 11, 12
    g = torch.Generator().manual_seed(5)  # same global problem on every rank
    X = torch.randn(N, D, dtype=torch.float64, generator=g)
    y = torch.sin(X.sum(-1, keepdim=True))
    Z = X[torch.randperm(N, generator=g)[:M]].clone() + 0.05
    k = cb.Matern52(variance=1.0, lengthscales=[1.5] * D)
    # single-GPU solve over ALL rows first (no communicator yet => no all-reduce), on every rank
    op1 = cb.SGPROperator(k, X.to(dev), Z.to(dev), 0.1)
    rhs1 = (op1.kuf_times(y.to(dev)) / 0.1).t().contiguous()
    sol1, (_, _, hist1) = cb.conjugate_gradient(op1, rhs1, None, 0.0, None, its, its + 1, return_history=True)
    # the same single-GPU solve through the independent two-sweep kernels: its distance to the fused path is the
    # rounding-noise floor of this (ill-conditioned) system, CG amplifies it from iteration to iteration
    op1s = cb.SGPROperator(k, X.to(dev), Z.to(dev), 0.1, variant=1)
    _, (_, _, hist1s) = cb.conjugate_gradient(op1s, rhs1, None, 0.0, None, its, its + 1, return_history=True)
    floor = torch.cummax(((hist1s - hist1).abs() / hist1).max(dim=1).values, dim=0).values
    # five right-hand sides (the probe solves of cggp/models.py:308-314): 8-wide DMMA-contraction sweeps
    g5 = torch.Generator().manual_seed(11)
    rhs5 = torch.randn(5, M, dtype=torch.float64, generator=g5).to(dev)
    sol5_1, (_, _, hist5_1) = cb.conjugate_gradient(op1, rhs5, None, 0.0, None, 6, 7, return_history=True)
    del op1, op1s
    # sharded solve: this rank's rows, one NCCL all-reduce per operator application
    ctx.init_comm()
    s, e = shard_rows(N, rank, world)
    op = cb.SGPROperator(k, X[s:e].to(dev), Z.to(dev), 0.1)
    rhs = (op.kuf_times(y[s:e].to(dev)) / 0.1).t().contiguous()
    sol, (steps, _, hist) = cb.conjugate_gradient(op, rhs, None, 0.0, None, its, its + 1, return_history=True)
    # every rank must hold bit-identical iterates (replicated scalars, no second collective)
    gathered = [torch.empty_like(sol) for _ in range(world)]
    dist.all_gather(gathered, sol)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    dh = ((hist - hist1).abs() / hist1).max(dim=1).values
    early = float(dh[:4].max())
    within = bool((dh <= torch.clamp(50.0 * floor, min=2e-9)).all())
    ok = same and early < 2e-9 and within and int(steps) == its
    sol5, (_, _, hist5) = cb.conjugate_gradient(op, rhs5, None, 0.0, None, 6, 7, return_history=True)
    g5l = [torch.empty_like(sol5) for _ in range(world)]
    dist.all_gather(g5l, sol5)
    same5 = all(torch.equal(g5l[0], t) for t in g5l)
    early5 = float(((hist5 - hist5_1).abs() / hist5_1)[:3].max())
    ok = ok and same5 and early5 < 2e-9
    # float32 leg: the tcgen05 TF32 kernels on the shards + the same all-reduce, against one GPU over all rows
    N32, M32, D32 = 60_000, 512, 40
    g32 = torch.Generator().manual_seed(9)
    X32 = torch.randn(N32, D32, dtype=torch.float32, generator=g32)
    Z32 = X32[torch.randperm(N32, generator=g32)[:M32]].clone()
    V32 = torch.randn(2, M32, dtype=torch.float32, generator=g32).to(dev)
    k32 = cb.SquaredExponential(1.0, [D32 ** 0.5] * D32)
    s32, e32 = shard_rows(N32, rank, world)
    op_sh = cb.SGPROperator(k32, X32[s32:e32].to(dev), Z32.to(dev), 0.1)
    w_sh = op_sh.kuf_kfu_matmul(V32)  # all-reduced
    op_all = cb.SGPROperator(k32, X32.to(dev), Z32.to(dev), 0.1)
    w_all = op_all.kuf_kfu_matmul(V32, allreduce=False)
    dev32 = float((w_sh - w_all).abs().max() / w_all.abs().max())
    ok = ok and op_sh.X32 is not None and dev32 < 1e-5
    if rank == 0:
        msg = (f"world={world} peer_tail={ctx.peer_allreduce} identical_on_ranks={same} early_dev={early:.2e} "
               f"max_dev={float(dh.max()):.2e} rhs5_identical={same5} rhs5_early_dev={early5:.2e} "
               f"tf32_sharded_vs_single={dev32:.2e}")
        print(("MGPU_CHECK ok " if ok else "MGPU_CHECK FAIL ") + msg, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
