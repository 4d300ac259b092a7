"""Short single-GPU cases for ncu (one kernel family per run, few launches; the same command line must first exit 0
without ncu - B200_PROFILING.md):  python tools/prof_case.py pipe1 | pipe8 | cg | gram [rows] [c3 | c2 | c4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cggp_b200 as cb


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "pipe1"
    g = torch.Generator(device="cuda").manual_seed(0)
    N, M, D = (int(sys.argv[2]) if len(sys.argv) > 2 else 500_000), 4096, 11
    shape = sys.argv[3] if len(sys.argv) > 3 else "c3"
    if shape == "c2":
        M, D = 2048, 3
    elif shape == "c4":
        M, D = 16384, 2
    X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
    Z = torch.randn(M, D, dtype=torch.float64, device="cuda", generator=g)
    k = (cb.SquaredExponential if shape == "c2" else cb.Matern52)(variance=1.0, lengthscales=[1.0] * D)
    op = cb.SGPROperator(k, X, Z, 0.1)
    if which in ("pipe1", "pipe8"):
        V = torch.randn(1 if which == "pipe1" else 8, M, dtype=torch.float64, device="cuda", generator=g)
        for _ in range(4):
            W = op.kuf_kfu_matmul(V, variant=3)
        torch.cuda.synchronize()
        print(which, float(W.abs().max()))
    elif which == "cg":
        rhs = torch.randn(1, M, dtype=torch.float64, device="cuda", generator=g)
        sol, (steps, _) = cb.conjugate_gradient(op, rhs, None, 0.0, None, 6, 7)
        torch.cuda.synchronize()
        print(which, int(steps), float(sol.abs().max()))
    elif which == "gram":
        G = op.gram(op.PX.rows(0, 65536))
        torch.cuda.synchronize()
        print(which, float(G.abs().max()))


if __name__ == "__main__":
    main()
