"""Device timing of the dense CDGP path (development aid): `V @ A` for B right-hand sides (cggp_symm_matmul: GEMV kernel for
B <= 8, DMMA tile GEMM above) against cuBLAS DGEMM, and a full multi-RHS CG solve on Kuu + Lambda (CGGP.predict_f's hot
loop, cggp/models.py:340)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cggp_b200 as cb


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for M in (500, 2048, 4096, 16384):
        A = torch.randn(M, M, dtype=torch.float64, device="cuda", generator=g)
        A = A @ A.t() / M + torch.eye(M, dtype=torch.float64, device="cuda")
        op = cb.DenseOperator(A)
        for B in (1, 5, 64, 1024, 4096):
            if M == 16384 and B > 1024:
                continue
            V = torch.randn(B, M, dtype=torch.float64, device="cuda", generator=g)
            t = timeit(lambda: op.matmul(V))
            tc = timeit(lambda: V @ A)
            err = float((op.matmul(V) - V @ A).abs().max() / (V @ A).abs().max())
            flop = 2.0 * B * M * M
            print(f"M={M:6d} B={B:5d}: symm_matmul {t:8.3f} ms ({flop / t / 1e9:7.2f} TFLOP/s, {M * M * 8 / t / 1e6:7.1f} GB/s of A)"
                  f"   cuBLAS {tc:8.3f} ms ({flop / tc / 1e9:7.2f} TFLOP/s)   rel diff {err:.1e}", flush=True)
    # CGGP.predict_f-shaped solve: M = 2048, B = 2000 right-hand sides
    M, B = 2048, 2000
    X = torch.randn(M, 3, dtype=torch.float64, device="cuda", generator=g)
    k = cb.SquaredExponential(1.0, [1.0] * 3)
    A = cb.add_diagonal(cb.Kuu(X, k), torch.full((M,), 0.01, dtype=torch.float64, device="cuda"))
    rhs = torch.randn(B, M, dtype=torch.float64, device="cuda", generator=g)
    its = 50
    t = timeit(lambda: cb.conjugate_gradient(A, rhs, None, 0.0, None, its, its + 1), reps=3, warm=1)
    print(f"dense CG M={M} B={B}: {its} iterations in {t:.2f} ms -> {t / its:.3f} ms/it, "
          f"{2.0 * B * M * M * its / t / 1e9:.2f} TFLOP/s on the product", flush=True)


if __name__ == "__main__":
    main()
