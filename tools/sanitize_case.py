"""Small invocations of every hand-written kernel family for compute-sanitizer (development aid):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cggp_b200 as cb
from cggp_b200 import selection


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for (N, M, D, kern) in ((1237, 300, 11, "matern52"), (999, 130, 3, "se"), (50, 700, 7, "matern32"),
                            (640, 270, 20, "matern52")):
        X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
        Z = torch.randn(M, D, dtype=torch.float64, device="cuda", generator=g)
        V = torch.randn(3, M, dtype=torch.float64, device="cuda", generator=g)
        y = torch.randn(N, 2, dtype=torch.float64, device="cuda", generator=g)
        k = cb.kernels.KERNELS[kern](1.1, [1.3] * D)
        op = cb.SGPROperator(k, X, Z, 0.1)
        ws = [op.kuf_kfu_matmul(V, variant=v) for v in (1, 3)]
        assert float((ws[1] - ws[0]).abs().max() / ws[0].abs().max()) < 1e-10
        # 8-wide DMMA-contraction kernel: full, ragged and two-sweep right-hand-side counts; Kuf @ Y with 5 columns
        for B8 in (8, 5, 11):
            V8 = torch.randn(B8, M, dtype=torch.float64, device="cuda", generator=g)
            w8, w1 = op.kuf_kfu_matmul(V8, variant=3), op.kuf_kfu_matmul(V8, variant=1)
            assert float((w8 - w1).abs().max() / w1.abs().max()) < 1e-10
        op.kuf_times(torch.randn(N, 5, dtype=torch.float64, device="cuda", generator=g))
        op.kuf_times(y)
        G = op.gram()
        assert float((G - G.t()).abs().max()) == 0.0
        rhs = torch.randn(2, M, dtype=torch.float64, device="cuda", generator=g)
        cb.conjugate_gradient(op, rhs, None, 0.0, None, 5, 3)
        pc = op.nystrom_preconditioner(num_rows=2 * M)
        cb.conjugate_gradient(op, rhs, None, 1e-8, pc, 20, 100)
        if M % 10 == 0:
            blocks = torch.randperm(M, generator=torch.Generator().manual_seed(1)).reshape(M // 10, 10)
            cb.BlockPreconditioner(blocks)(rhs, op.Kuu)
            cb.conjugate_gradient(op.Kuu, rhs, None, 1e-9, cb.BlockPreconditioner(blocks), 10, 100)
        with torch.no_grad():
            mdl = cb.cdgp_class(k, cb.Gaussian(0.1), Z, error_threshold=1e-8, pseudo_u=y[:M, :1].contiguous()
                                if N >= M else None)
            mdl.predict_f(X[:33])
            mdl.elbo((X[:40], y[:40, :1]))
        selection.nearest_center_update(Z, (X, y[:, :1]))
        selection.kmeans_indices_and_distances(Z, X)
        A = cb.add_diagonal(cb.Kuu(Z, k), torch.full((M,), 0.1, dtype=torch.float64, device="cuda"))
        for B in (1, 5, 40, 200):
            cb.conjugate_gradient(A, torch.randn(B, M, dtype=torch.float64, device="cuda", generator=g), None, 1e-9,
                                  None, 15, 7)
        Xf, Zf, Vf = X.float(), Z.float(), V.float()
        opf = cb.SGPROperator(cb.kernels.KERNELS[kern](1.1, [1.3] * D), Xf, Zf, 0.1)
        w4, w1 = opf.kuf_kfu_matmul(Vf), opf.kuf_kfu_matmul(Vf, variant=1)
        assert float((w4 - w1).abs().max() / w1.abs().max()) < 1e-3
        opf.kuf_times(y.float())
        var = torch.tensor(1.1, dtype=torch.float64, device="cuda", requires_grad=True)
        ls = torch.full((D,), 1.3, dtype=torch.float64, device="cuda", requires_grad=True)
        cb.kernels.KERNELS[kern](var, ls).K(X, Z).sum().backward()
    torch.cuda.synchronize()
    print("sanitize_case ok")


if __name__ == "__main__":
    main()
