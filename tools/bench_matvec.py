"""Quick device timing of the matrix-free product and one CG iteration (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cggp_b200 as cb

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)

def main():
    cfgs = [("c3", 2_000_000, 4096, 11, "matern52"), ("c2", 434_874, 2048, 3, "se"), ("c1", 10_000, 500, 2, "se"),
            ("c4", 1_000_000, 16384, 2, "matern52")]  # c4: one rank's share (N = 8M over 8 GPUs) of BASELINE configs[3]
    if len(sys.argv) == 1:
        cfgs = cfgs[:3]
    if len(sys.argv) > 1:
        cfgs = [c for c in cfgs if c[0] in sys.argv[1:]]
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, N, M, D, kern in cfgs:
        X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
        Z = torch.randn(M, D, dtype=torch.float64, device="cuda", generator=g)
        V = torch.randn(1, M, dtype=torch.float64, device="cuda", generator=g)
        k = cb.kernels.KERNELS[kern](variance=1.0, lengthscales=[1.0] * D)
        op = cb.SGPROperator(k, X, Z, 0.1)
        falg = 2.0 * N * M * (D + 2)
        for variant in (3,):
            best, avg = timeit(lambda: op.kuf_kfu_matmul(V, variant=variant))
            print(f"{name}: fused matvec v{variant} N={N} M={M} D={D} {kern}: best {best:.3f} ms avg {avg:.3f} ms  "
                  f"-> {falg / best / 1e9:.2f} TFLOP/s F_alg, {N * M / best / 1e6:.1f} Gentry/s", flush=True)
        if N <= 500_000:
            best1, _ = timeit(lambda: op.kuf_kfu_matmul(V, variant=1), reps=3, warm=1)
            print(f"{name}: simple matvec best {best1:.3f} ms ({best1 / best:.1f}x fused)")
        # full CG iterations (fixed count, threshold 0)
        rhs = torch.randn(1, M, dtype=torch.float64, device="cuda", generator=g)
        its = 20
        def solve():
            cb.conjugate_gradient(op, rhs, None, 0.0, None, its, its + 1)
        best, avg = timeit(solve, reps=3, warm=1)
        print(f"{name}: CG {its} its: {best:.2f} ms -> {its / best * 1e3:.1f} it/s")

def main_multi_rhs(names):
    """F_alg fraction of the fused product as a function of the number of right-hand sides (B = 1, 2: FMA contractions,
    matvec_pipe.cu; B >= 3: DMMA contractions, matvec_pipe8.cu).  F_alg(B) = 2 N M (D + 2 B), SURVEY.md 8(d)."""
    import ctypes as C
    from cggp_b200 import _lib
    ctx = _lib.context()
    gops = C.c_double(0.0)
    ctx.check(ctx.lib.cggp_microbench(ctx.handle, 1, 4096, C.byref(gops)))
    peak = gops.value / 1e3
    print(f"FP64 DMMA issue-rate peak measured here: {peak:.2f} TFLOP/s")
    cfgs = {"c3": (2_000_000, 4096, 11, "matern52"), "c2": (434_874, 2048, 3, "se"),
            "c4": (1_000_000, 16384, 2, "matern52"), "c3q": (500_000, 4096, 11, "matern52")}
    g = torch.Generator(device="cuda").manual_seed(0)
    for name in names:
        N, M, D, kern = cfgs[name]
        X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
        Z = torch.randn(M, D, dtype=torch.float64, device="cuda", generator=g)
        k = cb.kernels.KERNELS[kern](variance=1.0, lengthscales=[1.0] * D)
        op = cb.SGPROperator(k, X, Z, 0.1)
        V16 = torch.randn(16, M, dtype=torch.float64, device="cuda", generator=g)
        W1 = torch.cat([op.kuf_kfu_matmul(V16[b:b + 1], variant=3) for b in range(16)])
        for B in (1, 2, 3, 4, 8, 16):
            V = V16[:B].contiguous()
            best, avg = timeit(lambda: op.kuf_kfu_matmul(V, variant=3), reps=4, warm=2)
            W = op.kuf_kfu_matmul(V, variant=3)
            err = float((W - W1[:B]).abs().max() / W1[:B].abs().max())
            falg = 2.0 * N * M * (D + 2 * B)
            print(f"{name} B={B:2d}: best {best:8.3f} ms avg {avg:8.3f} ms  {falg / best / 1e9:6.2f} TFLOP/s F_alg = "
                  f"{falg / best / 1e9 / peak:.3f} of peak; {N * M / best / 1e6:.1f} Gentry/s; "
                  f"max rel diff vs single-RHS sweeps {err:.2e}", flush=True)


def main_f32():
    """config 5 shape (float32, D = 90, M = 8192): tensor-core path vs the FFMA two-sweep kernels, reduced N."""
    g = torch.Generator(device="cuda").manual_seed(0)
    N, M, D = 500_000, 8192, 90
    X = torch.randn(N, D, dtype=torch.float32, device="cuda", generator=g)
    Z = torch.randn(M, D, dtype=torch.float32, device="cuda", generator=g)
    V = torch.randn(1, M, dtype=torch.float32, device="cuda", generator=g)
    k = cb.SquaredExponential(1.0, [D ** 0.5] * D)
    for nsplit in (16, 3, 1):
        op = cb.SGPROperator(k, X, Z, 0.1, variant=4, tf32_nsplit=nsplit)
        best, avg = timeit(lambda: op.kuf_kfu_matmul(V), reps=3, warm=1)
        flop = 2.0 * 2.0 * N * M * 96 * (1 if nsplit == 1 else 3)
        name = "3xFP16" if nsplit == 16 else f"TF32 x{nsplit}"
        print(f"c5/4 (N={N}): tcgen05 {name}: best {best:.3f} ms -> {flop / best / 1e9:.1f} TFLOP/s tensor, "
              f"{2 * N * M / best / 1e6:.1f} Gentry/s (two sweeps)", flush=True)
        if nsplit == 16:
            W16 = op.kuf_kfu_matmul(V)
    op = cb.SGPROperator(k, X, Z, 0.1, variant=4, tf32_nsplit=3)
    W4 = op.kuf_kfu_matmul(V)
    print(f"c5/4: max rel diff 3xFP16 vs 3xTF32 {float((W16 - W4).abs().max() / W4.abs().max()):.3e}", flush=True)
    best1, _ = timeit(lambda: op.kuf_kfu_matmul(V, variant=1), reps=2, warm=1)
    W1 = op.kuf_kfu_matmul(V, variant=1)
    print(f"c5/4: FFMA two-sweep: best {best1:.3f} ms; max rel diff tf32x3 vs ffma "
          f"{float((W4 - W1).abs().max() / W1.abs().max()):.3e}", flush=True)


if __name__ == "__main__":
    if "multi" in sys.argv[1:]:
        main_multi_rhs([a for a in sys.argv[1:] if a != "multi"] or ["c3"])
    elif "c5" in sys.argv[1:]:
        main_f32()
    else:
        main()
