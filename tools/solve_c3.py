"""Seconds per FULL solve of the north-star system at a named workload (development aid; bench.py is the contract):
Nystrom-preconditioned matrix-free CG to the reference's absolute threshold 0.5|r|^2 <= 1e-6 (cggp/cli_utils.py:439)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cggp_b200 as cb

CFG = {"c3": (2_000_000, 4096, 11, "matern52"), "c2": (434_874, 2048, 3, "se"), "c1": (10_000, 500, 2, "se")}


def main():
    for name in (sys.argv[1:] or ["c3"]):
        N, M, D, kern = CFG[name]
        g = torch.Generator(device="cuda").manual_seed(0)
        X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
        y = torch.sin(X.sum(-1, keepdim=True)) + 0.3 * torch.randn(N, 1, dtype=torch.float64, device="cuda", generator=g)
        Z = X[torch.randperm(N, device="cuda", generator=g)[:M]].clone()
        k = cb.kernels.KERNELS[kern](variance=1.0, lengthscales=[1.0] * D)
        op = cb.SGPROperator(k, X, Z, 0.1)
        rhs = (op.kuf_times(y) / 0.1).t().contiguous()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pc = op.nystrom_preconditioner()
        torch.cuda.synchronize()
        t_pc = time.perf_counter() - t0
        for label, p, maxit in (("nystrom", pc, 300), ("plain", None, 100)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sol, (steps, err, h) = cb.conjugate_gradient(op, rhs, None, 1e-6, p, maxit, maxit + 1, return_history=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"{name} {label}: steps={int(steps)} time={dt:.3f}s (+{t_pc:.3f}s preconditioner setup) "
                  f"0.5|r|^2: start {float(h[0].max()):.3e} -> end {float(h[-1].max()):.3e}", flush=True)


if __name__ == "__main__":
    main()
