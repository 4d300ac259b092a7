"""Device timing of the point kernels (development aid): prepare_points, kernel matrix, nearest-centre assignment,
cluster statistics - the kernels either side of the CG hot path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cggp_b200 as cb
from cggp_b200 import selection


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
    N, M, D = 2_000_000, 4096, 11
    X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(N, 1, dtype=torch.float64, device="cuda", generator=g)
    Z = X[torch.randperm(N, device="cuda", generator=g)[:M]].clone()
    k = cb.Matern52(1.0, [1.0] * D)
    t = timeit(lambda: k.prepare(X))
    print(f"prepare_points N={N} D={D}: {t:.3f} ms -> {N * (D + 13) * 8 / t / 1e6:.0f} GB/s (read X, write P + norms)")
    PX, PZ = k.prepare(X), k.prepare(Z)
    n1 = 100_000
    sub = PX.rows(0, n1)
    out = torch.empty((n1, M), dtype=torch.float64, device="cuda")
    t = timeit(lambda: cb.kernels.kernel_matrix(k.kind, k.variance, sub, PZ, out=out))
    print(f"kernel_matrix {n1} x {M} D={D} matern52: {t:.3f} ms -> {n1 * M / t / 1e6:.1f} Gentry/s, "
          f"{n1 * M * 8 / t / 1e6:.0f} GB/s written")
    t = timeit(lambda: cb.kernels.kernel_matrix(k.kind, k.variance, PZ, PZ), reps=5)
    print(f"Kuu {M} x {M}: {t:.3f} ms")
    t = timeit(lambda: selection.nearest_center_update(Z, (X, y)), reps=3, warm=1)
    print(f"nearest_center_update N={N} M={M} D={D} (sq-euclidean argmin + cluster stats): {t:.3f} ms -> "
          f"{N * M / t / 1e6:.1f} Gpair/s, {2.0 * N * M * D / t / 1e9:.2f} TFLOP/s")
    idx, _ = selection._nearest(X, Z, "sqeuclidean", None)
    t = timeit(lambda: selection.cluster_stats(idx, y, M))
    print(f"cluster_stats N={N}: {t:.3f} ms -> {N * 16 / t / 1e6:.0f} GB/s")
    # float32 nearest centre at the config-5 shape (reduced N)
    N5, M5, D5 = 500_000, 8192, 90
    X5 = torch.randn(N5, D5, dtype=torch.float32, device="cuda", generator=g)
    Z5 = X5[:M5].clone()
    y5 = torch.randn(N5, 1, dtype=torch.float32, device="cuda", generator=g)
    t = timeit(lambda: selection.nearest_center_update(Z5, (X5, y5)), reps=2, warm=1)
    print(f"nearest_center_update float32 N={N5} M={M5} D={D5}: {t:.3f} ms -> {2.0 * N5 * M5 * D5 / t / 1e9:.2f} TFLOP/s")


if __name__ == "__main__":
    main()
