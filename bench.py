#!/usr/bin/env python
"""bench.py - the headline benchmark of BASELINE.json: float64 matrix-free CG iterations / s on the north-star
operator  Sigma = Kuu + s^-2 Kuf Kfu  at  N = 2 000 000, M = 4096, D = 11, Matern-5/2  (config c3), N sharded over
the GPUs of one box (one process per GPU, one NCCL all-reduce of the partial M-vector per iteration).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c3|c2|c1|c4|c5]

A "step" is ONE CG iteration (cggp/conjugate_gradient.py:64-85): one application of Sigma to the search direction
(fused Kuf Kfu product + all-reduce + Kuu product) and the fused vector update.  The timed region is one call of
``cggp_b200.conjugate_gradient(operator, rhs, None, 0.0, None, K, K + 1)`` (threshold 0 => exactly K iterations)
through the Python mirror of the reference API -> ctypes -> libcggp_b200.so.  Prints ONE JSON line on rank 0.

--impl reference times the reference's own CPU path (the torch-CPU port of the restated reference in oracle/, all
host threads) on a bounded row sample of the same workload, scaled linearly in N (every N-dependent cost of an
iteration is exactly linear in N); TensorFlow + GPflow are not installable in this image.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, M, D, kernel, description)
    "c3": (2_000_000, 4096, 11, "matern52", "CDGP/SGPR CG solve, synthetic houseelectric-shaped (BASELINE configs[2])"),
    "c2": (434_874, 2048, 3, "se", "synthetic 3droad-shaped (BASELINE configs[1])"),
    "c1": (10_000, 500, 2, "se", "synthetic 2-D regression (BASELINE configs[0])"),
    "c4": (8_000_000, 16384, 2, "matern32", "synthetic geospatial-shaped, SGPR system (BASELINE configs[3]; quoted on 8 B200)"),
    "c5": (2_000_000, 8192, 90, "se", "float32, YearPredictionMSD-shaped, tensor-core distance GEMM (BASELINE configs[4])"),
}
FLOAT32 = {"c5"}   # every other workload is float64
NOISE = 0.1        # likelihood variance, cggp/cli_utils.py:153
METRIC = "fp64 CG iter/s at N=2M,M=4096,D=11 on 1/2/4/8 B200; % of FP64 peak"  # BASELINE.json's metric (c3)


def workload_string(workload):
    """Identical in both arms (the driver compares the configs of the native and the reference line)."""
    N, M, D, kern, desc = WORKLOADS[workload]
    return f"{workload}: N={N}, M={M}, D={D}, {kern}, B=1 right-hand side; {desc}"


def l2_string(workload, world):
    """Timing rule: inputs larger than L2 (identical text in both arms; it describes the GPU arm's working set)."""
    N, M, D, _, _ = WORKLOADS[workload]
    esize = 4 if workload in FLOAT32 else 8
    n_local = (N + world - 1) // world
    return ("inputs larger than L2, no flush needed: per GPU the prepared X shard %.0f MB + Kuu %.0f MB are streamed "
            "every iteration (126 MB L2)" % (n_local * (4 * ((D + 4) // 4) + 2) * esize / 1e6, M * M * esize / 1e6))


OPERATOR_STRING = "Kuu + jitter I + Kuf Kfu / noise_variance, matrix-free (Kfu never materialised)"
STEP_STRING = "one CG iteration: Kuf Kfu product (+ all-reduce over ranks) + Kuu product + vector update"


def metric_for(workload):
    if workload == "c3":
        return METRIC
    N, M, D, kern, _ = WORKLOADS[workload]
    return f"{'fp32' if workload in FLOAT32 else 'fp64'} CG iter/s at N={N},M={M},D={D} ({kern}); not BASELINE.json's headline config"


def f_alg_matvec(n_rows, m, d, b=1):
    """Algorithmic flop of ONE fused Kuf Kfu product launch: distance contraction 2 N M D plus the
    two tile contractions 2 * 2 N M B; the element-wise kernel epilogue (sqrt, exp, polynomial) is NOT counted."""
    return 2.0 * n_rows * m * (d + 2 * b)


def f_alg_iteration(n_rows, m, d, b=1):
    return f_alg_matvec(n_rows, m, d, b) + 2.0 * m * m * b


# ----------------------------------------------------------------------------------------------------------------
# clocks / throttle reasons sampled DURING the timed region (NVML; nvidia-smi as a fallback)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}

    def __init__(self, device_index: int, period: float = 0.02):
        self.idx, self.period = device_index, period
        self.samples, self.reasons, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = device_index
            if vis:
                toks = [t for t in vis.split(",") if t.strip() != ""]
                if device_index < len(toks) and toks[device_index].strip().isdigit():
                    phys = int(toks[device_index])
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        names = [k for k, bit in {**self.BAD, **self.NOTE}.items() if self.reasons & bit]
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": names, "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the restated reference (oracle/torch_cpu.py) on the host cores, bounded sample, scaled linearly in N
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_arm(workload: str, steps: int, warmup: int, budget_s: float):
    import torch

    from oracle import torch_cpu as tc  # checker / baseline only: the one place bench.py executes oracle/

    N, M, D, kern, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(7)
    dt = torch.float32 if workload in FLOAT32 else torch.float64
    ls = torch.full((D,), math.sqrt(D) if workload in FLOAT32 else 1.0, dtype=dt)
    Z = torch.randn(M, D, dtype=dt, generator=g)
    rhs = torch.randn(1, M, dtype=dt, generator=g)

    def time_iters(n_rows, iters, warm):
        X = torch.randn(n_rows, D, dtype=dt, generator=g)
        mm = tc.sgpr_operator(kern, 1.0, ls, X, Z, NOISE, chunk=8192)
        if warm:
            tc.cg_iterations(mm, rhs, warm)
        t0 = time.perf_counter()
        tc.cg_iterations(mm, rhs, iters)
        return (time.perf_counter() - t0) / iters

    probe_rows = min(N, 8192)
    t_probe = time_iters(probe_rows, 1, 1)
    per_row = t_probe / probe_rows
    total_iters = steps + warmup
    rows = int(min(N, max(probe_rows, budget_s / max(total_iters, 1) / max(per_row, 1e-12))))
    rows = max(1024, (rows // 1024) * 1024) if rows < N else N
    t_iter_sample = time_iters(rows, steps, warmup)
    # N-linear part scaled to the full N; the M^2 part (Kuu product) is identical and tiny (kept inside the sample)
    t_iter_full = t_iter_sample * (N / rows)
    return {
        "value": 1.0 / t_iter_full,
        "unit": "CG iterations/s",
        "cores": cores,
        "kind": "port",
        "sample": f"{rows} of {N} rows (same M={M}, D={D}, {kern}, {str(dt)[6:]}), {steps} timed CG iterations after "
                  f"{warmup} warm-up, per-iteration time scaled by N/rows = {N / rows:.2f} (cost is linear in N)",
        "ms_per_step_sample": t_iter_sample * 1e3,
        "ms_per_step_scaled": t_iter_full * 1e3,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 times the CPU path
    N, M, D, kern, desc = WORKLOADS[args.workload]
    base = cpu_reference_arm(args.workload, args.steps, args.warmup, budget_s=90.0)
    line = {
        "impl": "reference",
        "metric": metric_for(args.workload), "value": base["value"], "unit": "CG iterations/s", "n_gpus": args.gpus,
        # a step of this arm is ONE CG iteration on the bounded row sample (what the clock saw); `value` is the rate of
        # the full workload, sample time x N / rows (every N-dependent cost of an iteration is exactly linear in N)
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step_sample"],
        "ms_per_step_full_workload_scaled": base["ms_per_step_scaled"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if args.workload in FLOAT32 else "f64", "data": "synthetic",
        # `config` is identical in both arms (the driver compares them); arm-specific detail goes to `run`
        "config": {"workload": workload_string(args.workload), "operator": OPERATOR_STRING, "step": STEP_STRING,
                   "l2": l2_string(args.workload, max(args.gpus, 1))},
        "run": {"impl_note": "restated reference (torch-CPU port of oracle/, chunked matrix-free product) on the host "
                             "cores, bounded row sample scaled linearly to the full N; n_gpus is echoed from the "
                             "command line; TensorFlow / GPflow are not installable in this image"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "CG iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# roofline denominators measured in the same run (MEASURED_PEAKS.json has HBM + BF16 only)
# ----------------------------------------------------------------------------------------------------------------
def fp64_peaks(ctx, device):
    """FP64 DMMA / DFMA issue rates, the throughput of the exp and sqrt routines the pipelined kernels use
    (cggp_microbench) and a cuBLAS DGEMM (torch.matmul float64, the denominator BASELINE.md 2 names) - library GEMM
    used ONLY as a yardstick."""
    import ctypes as C

    import torch

    out = {}
    for key, which in (("dmma_tflops", 1), ("dfma_tflops", 0), ("exp_gevals", 5), ("sqrt_gevals", 6)):
        g = C.c_double(0.0)
        try:
            ctx.check(ctx.lib.cggp_microbench(ctx.handle, which, 4096, C.byref(g)))
            out[key] = g.value / 1e3 if key.endswith("tflops") else g.value
        except Exception as exc:  # pragma: no cover
            out[key] = None
            out[key + "_error"] = str(exc)
    try:
        n = 6144
        A = torch.randn(n, n, device=device, dtype=torch.float64)
        B = torch.randn(n, n, device=device, dtype=torch.float64)
        for _ in range(2):
            A @ B
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            A @ B
        t1.record()
        torch.cuda.synchronize()
        out["dgemm_cublas_tflops"] = 5 * 2.0 * n ** 3 / (t0.elapsed_time(t1) * 1e-3) / 1e12
        del A, B
    except Exception as exc:  # pragma: no cover
        out["dgemm_cublas_tflops"] = None
        out["dgemm_error"] = str(exc)
    return out


def parity_single_rank(cb, device, its=12):
    """First half of the N > 1 parity check, run BEFORE the communicator exists: every rank solves the same 200k-row
    problem on its own GPU (no all-reduce) and keeps the residual history."""
    import torch

    N, M, D = 200_003, 1024, 11
    g = torch.Generator().manual_seed(5)
    X = torch.randn(N, D, dtype=torch.float64, generator=g)
    y = torch.sin(X.sum(-1, keepdim=True))
    Z = X[torch.randperm(N, generator=g)[:M]].clone() + 0.05
    k = cb.Matern52(variance=1.0, lengthscales=[1.5] * D)
    op1 = cb.SGPROperator(k, X.to(device), Z.to(device), 0.1)
    rhs1 = (op1.kuf_times(y.to(device)) / 0.1).t().contiguous()
    _, (_, _, h1) = cb.conjugate_gradient(op1, rhs1, None, 0.0, None, its, its + 1, return_history=True)
    # the same solve through the independent two-sweep kernels: rounding-noise floor of this ill-conditioned system
    op1s = cb.SGPROperator(k, X.to(device), Z.to(device), 0.1, variant=1)
    _, (_, _, h1s) = cb.conjugate_gradient(op1s, rhs1, None, 0.0, None, its, its + 1, return_history=True)
    floor = torch.cummax(((h1s - h1).abs() / h1).max(dim=1).values, dim=0).values
    return {"X": X, "y": y, "Z": Z, "kernel": k, "hist": h1, "floor": floor, "its": its}


def parity_sharded(cb, device, st, rank, world):
    """Second half (communicator initialised): the same problem with its rows sharded over the ranks through the code
    path the benchmark times; all ranks must hold bit-identical iterates, and the residual history must follow the
    1-rank solve (2e-9 on 0.5|r|^2 for the first iterations, then the measured rounding-noise floor x 50)."""
    import torch
    import torch.distributed as dist

    from cggp_b200.sharding import shard_rows

    s, e = shard_rows(st["X"].shape[0], rank, world)
    op = cb.SGPROperator(st["kernel"], st["X"][s:e].to(device), st["Z"].to(device), 0.1)
    rhs = (op.kuf_times(st["y"][s:e].to(device)) / 0.1).t().contiguous()
    its = st["its"]
    sol, (steps, _, h) = cb.conjugate_gradient(op, rhs, None, 0.0, None, its, its + 1, return_history=True)
    gathered = [torch.empty_like(sol) for _ in range(world)]
    dist.all_gather(gathered, sol)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    dh = ((h - st["hist"]).abs() / st["hist"]).max(dim=1).values
    early = float(dh[:4].max())
    within = bool((dh <= torch.clamp(50.0 * st["floor"], min=2e-9)).all())
    return {"problem": "N=200003, M=1024, D=11, matern52, 12 iterations, rows sharded as in the timed run",
            "rank_identical_iterates": bool(same), "early_rel_dev_vs_1rank": early,
            "max_rel_dev_vs_1rank": float(dh.max()), "noise_floor_last": float(st["floor"][-1]),
            "within_tolerance": bool(within and early < 2e-9), "steps": int(steps),
            "tolerance": "0.5|r|^2: 2e-9 relative on iterations 0-3, then max(2e-9, 50 x oracle-style noise floor)"}


def quick_its(cb, device, workload, steps=5):
    """CG iterations/s of another BASELINE config on this GPU (secondary figures of the c3 line)."""
    import torch

    N, M, D, kern, _ = WORKLOADS[workload]
    dt = torch.float32 if workload in FLOAT32 else torch.float64
    g = torch.Generator(device=device).manual_seed(99)
    X = torch.randn(N, D, dtype=dt, device=device, generator=g)
    Z = X[torch.randperm(N, device=device, generator=g)[:M]].clone()
    kernel = cb.kernels.KERNELS[kern](variance=1.0, lengthscales=[math.sqrt(D) if workload in FLOAT32 else 1.0] * D)
    op = cb.SGPROperator(kernel, X, Z, NOISE)
    rhs = torch.randn(1, M, dtype=dt, device=device, generator=g)
    cb.conjugate_gradient(op, rhs, None, 0.0, None, 3, 4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cb.conjugate_gradient(op, rhs, None, 0.0, None, steps, steps + 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    del op, X, Z
    torch.cuda.empty_cache()
    return {"value": steps / (ms * 1e-3), "unit": "CG iterations/s", "steps": steps, "ms_per_step": ms / steps,
            "workload": f"{workload}: N={N}, M={M}, D={D}, {kern}, {'f32' if workload in FLOAT32 else 'f64'}, B=1, 1 GPU"}


# ----------------------------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    import cggp_b200 as cb
    from cggp_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback on the product path"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    ctx = _lib.context(device)
    parity = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        pst = parity_single_rank(cb, device) if not args.no_parity_check else None
        ctx.init_comm()
        if pst is not None:
            parity = parity_sharded(cb, device, pst, rank, world)
            del pst
            torch.cuda.empty_cache()

    N, M, D, kern, desc = WORKLOADS[args.workload]
    from cggp_b200.sharding import shard_rows

    r_lo, r_hi = shard_rows(N, rank, world)
    n_local = r_hi - r_lo
    f32w = args.workload in FLOAT32
    f64 = torch.float32 if f32w else torch.float64  # the workload's dtype
    esize = 4 if f32w else 8

    # ---- synthetic data: HOST (pinned) copies for the e2e leg, device copies for the resident leg -------------
    kernel = cb.kernels.KERNELS[kern](variance=1.0, lengthscales=[math.sqrt(D) if f32w else 1.0] * D)

    def make_problem(rows):
        g = torch.Generator().manual_seed(1234 + rank)
        Xh_ = torch.randn(rows, D, dtype=f64, generator=g).pin_memory()
        yh_ = (torch.sin(Xh_.sum(-1, keepdim=True)) + math.sqrt(NOISE) * torch.randn(rows, 1, dtype=f64, generator=g))
        yh_ = yh_.pin_memory()
        # inducing points: M rows drawn from rank 0's shard (uniform without replacement, cggp/cli_utils.py:157-161)
        if rank == 0:
            sel = torch.randperm(rows, generator=g)[:M]
            Zh_ = Xh_[sel].clone()
        else:
            Zh_ = torch.empty(M, D, dtype=f64)
        Zd_ = Zh_.to(device)
        if world > 1:
            dist.broadcast(Zd_, src=0)
            Zh_ = Zd_.cpu()
        Zh_ = Zh_.pin_memory()
        Xd_, yd_ = Xh_.to(device), yh_.to(device)
        op_ = cb.SGPROperator(kernel, Xd_, Zd_, NOISE)
        rhs_ = (op_.kuf_times(yd_) / NOISE).t().contiguous()  # [1, M]  s^-2 Kuf y (all-reduced)
        return Xh_, yh_, Zh_, Zd_, Xd_, yd_, op_, rhs_

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve(operator, b, iters):
        return cb.conjugate_gradient(operator, b, None, 0.0, None, iters, iters + 1)

    Xh, yh, Zh, Zd, Xd, yd, op, rhs = make_problem(n_local)
    balance = None
    if args.balance and world > 1:
        # speed-weighted row split (opt-in): the devices of a box differ by a constant 1 - 2 %; calibrate the per-rank
        # product time on the even split, re-shard once in proportion to the measured rates (cggp_b200.sharding)
        from cggp_b200.sharding import shard_rows_weighted

        solve(op, rhs, 3)
        barrier()
        ctx.profile(True)
        solve(op, rhs, 12)
        torch.cuda.synchronize()
        pm = ctx.profile_read()["kuf_kfu_matvec"]
        ctx.profile(False)
        mine = torch.tensor([pm[0] / max(pm[1], 1) / max(n_local, 1)], dtype=torch.float64, device=device)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_row = [float(v.item()) for v in allr]
        r_lo, r_hi = shard_rows_weighted(N, rank, [1.0 / v for v in per_row])
        del op, Xd, yd, rhs
        torch.cuda.empty_cache()
        n_local = r_hi - r_lo
        Xh, yh, Zh, Zd, Xd, yd, op, rhs = make_problem(n_local)
        balance = {"ms_per_million_rows_even_split": [v * 1e6 for v in per_row], "rows_this_rank": n_local}

    # ---- resident leg: W warm-up iterations, then exactly K timed iterations ---------------------------------
    solve(op, rhs, max(args.warmup, 3))
    barrier()
    launches0 = ctx.launches
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        sol, (steps_done, half_rz) = solve(op, rhs, args.steps)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - launches0
    assert int(steps_done) == args.steps, (int(steps_done), args.steps)
    assert bool(torch.isfinite(sol).all())
    t = torch.tensor([ms], dtype=f64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- seconds per FULL solve (north star): Nystrom-preconditioned CG to the reference's absolute threshold -------
    full_solve = None
    if args.workload == "c3" and world == 1 and not args.no_secondary:
        try:
            f0, f1, f2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            f0.record()
            pc = op.nystrom_preconditioner()
            f1.record()
            fsol, (fsteps, ferr, fhist) = cb.conjugate_gradient(op, rhs, None, 1e-6, pc, 400, 401, return_history=True)
            f2.record()
            torch.cuda.synchronize()
            full_solve = {
                "what": "matrix-free CG on Sigma to 0.5|r|^2 <= 1e-6 (the reference's default error_threshold, "
                        "cggp/cli_utils.py:439), Nystrom preconditioner from a 4M-row subsample (plain CG stalls: "
                        "cond(Sigma) ~ 1e9+)",
                "iterations": int(fsteps), "seconds": f1.elapsed_time(f2) * 1e-3,
                "preconditioner_setup_seconds": f0.elapsed_time(f1) * 1e-3,
                "half_rr_start": float(fhist[0].max()), "half_rr_end": float(fhist[-1].max())}
            del pc, fsol
        except Exception as exc:  # pragma: no cover
            full_solve = {"error": str(exc)}

    # ---- e2e leg: same call, HOST buffers in, host result out, every copy inside the timed region -------------
    def e2e_once():
        Xe = Xh.to(device, non_blocking=True)
        ye = yh.to(device, non_blocking=True)
        Ze = Zh.to(device, non_blocking=True)
        ope = cb.SGPROperator(kernel, Xe, Ze, NOISE)
        be = (ope.kuf_times(ye) / NOISE).t().contiguous()
        s, (st, hz) = solve(ope, be, args.steps)
        out = torch.empty(s.shape, dtype=f64).pin_memory()
        out.copy_(s, non_blocking=True)
        torch.cuda.synchronize()
        return out

    del op, Xd, yd
    torch.cuda.empty_cache()
    e2e_once()  # warm-up (allocator, pinned staging)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    out = e2e_once()
    a1.record()
    barrier()
    t2 = torch.tensor([a0.elapsed_time(a1)], dtype=f64, device=device)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item())
    h2d = (Xh.numel() + yh.numel() + Zh.numel()) * esize
    d2h = out.numel() * esize

    # ---- roofline of the dominant kernel (the fused Kuf Kfu product), timed live with CUDA events ---------------
    mv_ms, mv_cnt = prof["kuf_kfu_matvec"]
    t_launch = (mv_ms / mv_cnt * 1e-3) if mv_cnt else None
    peak_tflops, peak_src, peaks = None, None, {}
    if f32w:
        # the float32 product runs 3xFP16 tcgen05 MMAs: the peak is the dense 16-bit tensor rate, sustained
        # (driver-measured in MEASURED_PEAKS.json; else a library GEMM timed here only as the denominator)
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak_tflops = float(mp["bf16_tflops_sustained"])
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (dense 16-bit tensor rate inside a long step)"
        except Exception:
            A_ = torch.randn(8192, 8192, device=device, dtype=torch.float16)
            B_ = torch.randn(8192, 8192, device=device, dtype=torch.float16)
            for _ in range(2):
                A_ @ B_
            t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0_.record()
            for _ in range(10):
                A_ @ B_
            t1_.record()
            torch.cuda.synchronize()
            peak_tflops = 10 * 2 * 8192 ** 3 / (t0_.elapsed_time(t1_) * 1e-3) / 1e12
            del A_, B_
            peak_src = "measured in this run: cuBLAS FP16 GEMM 8192^3 (torch.matmul); MEASURED_PEAKS.json absent"
    else:
        peaks = fp64_peaks(ctx, device)
        peak_tflops = peaks.get("dmma_tflops")
        peak_src = ("measured in this run: FP64 DMMA m8n8k4 issue-rate micro-benchmark (cggp_microbench; DFMA shares the "
                    "pipe at the same rate); cuBLAS DGEMM of the same run in peak_dgemm_cublas; MEASURED_PEAKS.json "
                    "has no FP64 figure")
    achieved = f_alg_matvec(n_local, M, D) / t_launch / 1e12 if t_launch else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    matern = kern.startswith("matern")
    roofline = {
        "kernel": ("tf32::gram_contract_kernel x2 (tcgen05 3xFP16 gram contraction, FP32 accumulate, two sweeps)" if f32w else
                   "kpipe::kfu_pipe_kernel (fused, software-pipelined Kuf Kfu product)"),
        "bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": (achieved / peak_tflops) if (achieved and peak_tflops) else None,
        "traffic": traffic, "peak_source": peak_src,
        "launches_timed": mv_cnt, "avg_launch_ms": (mv_ms / mv_cnt) if mv_cnt else None,
        "share_of_step": (mv_ms / ms) if ms > 0 else None,
        "sections_ms": {k: round(v[0], 4) for k, v in prof.items()},
        "gentries_per_s": (n_local * M / t_launch / 1e9) if t_launch else None,
    }
    if not f32w and t_launch:
        # report achieved_fp64 AND achieved_exp; the FP64 pipe is shared by the DMMA distance, the
        # exp / sqrt of every entry and the two contractions, so the bound is the SUM of their times at the measured
        # rates of each part running alone (composite), not any one of them
        E = float(n_local) * M
        pe, psq, pf = peaks.get("exp_gevals"), peaks.get("sqrt_gevals"), peaks.get("dfma_tflops")
        roofline["bound_detail"] = ("FP64 pipe, shared by the DMMA distance contraction and the FP64 exp"
                                    + (" + sqrt" if matern else "") + " of every Gram entry"
                                    + ("; exp dominates (one DMMA k-step)" if D <= 3 else ""))
        roofline["achieved_exp"] = E / t_launch / 1e9
        roofline["peak_exp"] = pe
        roofline["unit_exp"] = "G exp evaluations/s (the kernel's own routine, micro-benchmarked alone)"
        roofline["frac_exp"] = (E / t_launch / 1e9 / pe) if pe else None
        if matern:
            roofline["achieved_sqrt"] = E / t_launch / 1e9
            roofline["peak_sqrt"] = psq
            roofline["frac_sqrt"] = (E / t_launch / 1e9 / psq) if psq else None
        roofline["peak_dgemm_cublas"] = peaks.get("dgemm_cublas_tflops")
        roofline["frac_vs_dgemm_cublas"] = (achieved / peaks["dgemm_cublas_tflops"]) \
            if peaks.get("dgemm_cublas_tflops") else None
        roofline["peak_dfma"] = pf
        if peak_tflops and pe and pf and (psq or not matern):
            ks4 = 4 * ((D + 4) // 4)
            t_min = (2.0 * E * ks4 / (peak_tflops * 1e12) + E / (pe * 1e9) + (E / (psq * 1e9) if matern else 0.0)
                     + 2.0 * 2.0 * E / (pf * 1e12) + (3.0 if matern else 0.0) * 2.0 * E / (pf * 1e12) / 2.0)
            roofline["composite_fp64_pipe"] = {
                "t_min_ms": t_min * 1e3, "frac": t_min / t_launch,
                "note": "lower bound of one launch with every part at its stand-alone rate on the shared FP64 pipe: "
                        "padded distance DMMA + exp" + (" + sqrt + Matern polynomial" if matern else "")
                        + " + two contractions; frac = t_min / measured launch time"}
        # what the FP64 pipe actually executes per Gram entry (SASS count of the tile loop)
        slots = 4 * ((D + 4) // 4) + (17.5 if matern else 9.5)
        roofline["executed"] = {"fp64_fma_slots_per_entry": slots,
                                "frac_of_peak": achieved / peak_tflops * slots / (D + 2) if peak_tflops else None,
                                "note": "FP64-pipe occupancy by executed FMA slots; `frac` counts algorithmic flop only"}

    secondary = None
    if args.workload == "c3" and world == 1 and not args.no_secondary:
        secondary = {"c3_full_solve": full_solve}
        for wl in ("c2", "c5"):
            try:
                secondary[wl] = quick_its(cb, device, wl)
            except Exception as exc:  # pragma: no cover
                secondary[wl] = {"error": str(exc)}
        # BASELINE configs[1] as worded (cover-tree-selected inducing points -> CDGP ELBO -> predict_f), stage by stage
        try:
            import importlib.util

            spec = importlib.util.spec_from_file_location("config2_pipeline",
                                                          os.path.join(ROOT, "tools", "config2_pipeline.py"))
            config2_pipeline = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(config2_pipeline)
            secondary["c2_cdgp_pipeline"] = config2_pipeline.run(cb, device)
            torch.cuda.empty_cache()
        except Exception as exc:  # pragma: no cover
            secondary["c2_cdgp_pipeline"] = {"error": str(exc)}
        # the multi-RHS form of the headline product: 8 right-hand sides per sweep, both contractions on DMMA
        try:
            g8 = torch.Generator(device=device).manual_seed(3)
            X8 = torch.randn(min(N, 1_000_000), D, dtype=f64, device=device, generator=g8)
            op8 = cb.SGPROperator(kernel, X8, Zd, NOISE)
            V8 = torch.randn(8, M, dtype=f64, device=device, generator=g8)
            for _ in range(2):
                op8.kuf_kfu_matmul(V8)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for _ in range(3):
                op8.kuf_kfu_matmul(V8)
            b1.record()
            torch.cuda.synchronize()
            t8 = b0.elapsed_time(b1) / 3 * 1e-3
            a8 = f_alg_matvec(X8.shape[0], M, D, 8) / t8 / 1e12
            secondary["c3_b8_product"] = {
                "kernel": "kpipe8::kfu_pipe8_kernel (8 right-hand sides per sweep, t = K V^T and w = K^T t on DMMA)",
                "rows": int(X8.shape[0]), "ms_per_launch": t8 * 1e3, "achieved": a8, "unit": "TFLOP/s",
                "frac": a8 / peak_tflops if peak_tflops else None, "f_alg": "2 N M (D + 2 B), B = 8"}
            del op8, X8, V8
            torch.cuda.empty_cache()
        except Exception as exc:  # pragma: no cover
            secondary["c3_b8_product"] = {"error": str(exc)}

    hbm = None
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        pass

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            c = cpu_reference_arm(args.workload, 3, 1, budget_s=20.0)
            cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
        its = args.steps / (ms_max * 1e-3)
        line = {
            "metric": metric_for(args.workload), "value": its, "unit": "CG iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if f32w else "f64", "data": "synthetic",
            # `config` is identical in both arms (the driver compares them); the details of this run are in `run`
            "config": {"workload": workload_string(args.workload), "operator": OPERATOR_STRING, "step": STEP_STRING,
                       "l2": l2_string(args.workload, world)},
            "run": {
                "sharding": f"rows sharded over {world} GPU(s), {n_local} on rank 0; Z, Kuu and the CG vectors replicated"
                            + ("; speed-weighted split from a calibration solve (--balance)" if balance else ""),
                "balance": balance,
                "allreduce": (("fused tail kernel: rank-ordered sum over NVLink peer memory (incl. each rank's 1/W share "
                               "of p Kuu) + CG vector update in one launch" if ctx.peer_allreduce
                               else "ncclAllReduce, then the fused combine + vector-update kernel")
                              if world > 1 else "none (one rank); combine + vector update fused in one kernel"),
                "seconds_per_solve": f"{ms_max * 1e-3:.4f} s for {args.steps} iterations (threshold 0, fixed count)",
                "f_alg_per_iteration": f_alg_iteration(N, M, D),
                "fp64_frac_whole_iteration": (f_alg_iteration(N, M, D) * its / world / 1e12 / peak_tflops)
                if peak_tflops else None,
                "hbm_gbs_measured": hbm,
            },
            "roofline": roofline,
            "secondary": secondary,
            "parity_check": parity,
            "cpu_baseline": cpu,
            "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": "CG iterations/s",
                    "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                    "note": "one solve call from pinned HOST buffers: H2D of X shard, y, Z + point preparation + "
                            "Kuu + rhs + K iterations + D2H of the solution; bytes are per-solve totals / K"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50,
                    help="timed CG iterations (a full preconditioned solve of c3 takes ~107; see secondary.c3_full_solve)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--balance", action="store_true",
                    help="N > 1: speed-weighted row split from a calibration solve instead of the even one (opt-in)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the c2 / c5 / 8-RHS secondary figures (N = 1, c3)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the sharded-vs-1-rank parity check (N > 1)")
    ap.add_argument("--mode", default="cg", choices=["cg", "predict"],
                    help="cg: CG iterations/s (the headline); predict: SGPR vs CDGP predict_f (BASELINE configs[3])")
    ap.add_argument("--predict-points", type=int, default=100_000, help="held-out test points (all ranks together)")
    ap.add_argument("--predict-batch", type=int, default=4096, help="test points per predict_f call")
    ap.add_argument("--predict-rows", type=int, default=0, help="override N (reduced replicas of c4)")
    ap.add_argument("--predict-inducing", type=int, default=0, help="override M (rounded down to a square)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "predict":
        from tools.predict_bench import run_predict

        run_predict(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
