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


def step_kernel_bandwidth():
    """The fused CG vector update alone (cggp_cg_fused_step): reads p, pA, v, r and writes v, r, p = 7 B M sizeof bytes
    per call against the measured HBM copy bandwidth of MEASURED_PEAKS.json (6551.7 GB/s)."""
    import ctypes as C

    from cggp_b200 import _lib

    ctx = _lib.context()
    g = torch.Generator(device="cuda").manual_seed(1)
    for B, M in ((1, 4096), (5, 16384), (2000, 2048), (4096, 4096), (4096, 16384)):
        bufs = [torch.randn(B, M, dtype=torch.float64, device="cuda", generator=g) for _ in range(4)]
        pA, v, r, p = bufs
        rz = torch.ones(B, dtype=torch.float64, device="cuda")
        hrr = torch.empty(B, dtype=torch.float64, device="cuda")

        def call():
            ctx.use_current_stream()
            ctx.check(ctx.lib.cggp_cg_fused_step(ctx.handle, _lib.F64, B, M, _lib.ptr(pA), _lib.ptr(v), _lib.ptr(r),
                                                 _lib.ptr(p), _lib.ptr(rz), _lib.ptr(hrr), None))
        t = timeit(call, reps=5, warm=2)
        byts = 7.0 * B * M * 8
        print(f"cg_step_kernel B={B:5d} M={M:6d}: {t * 1e3:9.1f} us, {byts / 1e6:9.1f} MB -> {byts / t / 1e6:8.1f} GB/s "
              f"= {byts / t / 1e6 / 6551.7:.3f} of the measured HBM copy bandwidth", flush=True)


if __name__ == "__main__":
    if "step" in sys.argv[1:]:
        step_kernel_bandwidth()
    else:
        main()
