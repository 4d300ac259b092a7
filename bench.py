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
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step_scaled"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if args.workload in FLOAT32 else "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: N={N}, M={M}, D={D}, {kern}, B=1 right-hand side; {desc}",
                   "operator": "Kuu + jitter I + Kuf Kfu / noise_variance (matrix-free, chunked)",
                   "note": "CPU path uses host cores only; n_gpus is echoed from the command line"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "CG iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        ctx.init_comm()

    N, M, D, kern, desc = WORKLOADS[args.workload]
    from cggp_b200.sharding import shard_rows

    r_lo, r_hi = shard_rows(N, rank, world)
    n_local = r_hi - r_lo
    f32w = args.workload in FLOAT32
    f64 = torch.float32 if f32w else torch.float64  # the workload's dtype
    esize = 4 if f32w else 8

    # ---- synthetic data: HOST (pinned) copies for the e2e leg, device copies for the resident leg -------------
    g = torch.Generator().manual_seed(1234 + rank)
    Xh = torch.randn(n_local, D, dtype=f64, generator=g).pin_memory()
    yh = (torch.sin(Xh.sum(-1, keepdim=True)) + math.sqrt(NOISE) * torch.randn(n_local, 1, dtype=f64, generator=g))
    yh = yh.pin_memory()
    # inducing points: M rows drawn from rank 0's shard (uniform without replacement, cggp/cli_utils.py:157-161)
    if rank == 0:
        sel = torch.randperm(n_local, generator=g)[:M]
        Zh = Xh[sel].clone()
    else:
        Zh = torch.empty(M, D, dtype=f64)
    Zd = Zh.to(device)
    if world > 1:
        dist.broadcast(Zd, src=0)
        Zh = Zd.cpu()
    Zh = Zh.pin_memory()
    kernel = cb.kernels.KERNELS[kern](variance=1.0, lengthscales=[math.sqrt(D) if f32w else 1.0] * D)

    Xd, yd = Xh.to(device), yh.to(device)
    op = cb.SGPROperator(kernel, Xd, Zd, NOISE)
    rhs = (op.kuf_times(yd) / NOISE).t().contiguous()  # [1, M]  s^-2 Kuf y (all-reduced)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve(operator, b, iters):
        return cb.conjugate_gradient(operator, b, None, 0.0, None, iters, iters + 1)

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
    peak_tflops, peak_src = None, None
    try:
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
            import ctypes as C

            gops = C.c_double(0.0)
            ctx.check(ctx.lib.cggp_microbench(ctx.handle, 1, 4096, C.byref(gops)))
            peak_tflops = gops.value / 1e3
            peak_src = ("measured in this run: FP64 DMMA m8n8k4 issue-rate micro-benchmark (cggp_microbench); "
                        "MEASURED_PEAKS.json has no FP64 figure")
    except Exception as exc:  # pragma: no cover
        peak_src = f"unavailable: {exc}"
    achieved = f_alg_matvec(n_local, M, D) / (mv_ms / max(mv_cnt, 1) * 1e-3) / 1e12 if mv_cnt else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "kernel": ("tf32::gram_contract_kernel x2 (tcgen05 3xFP16 gram contraction, FP32 accumulate, two sweeps)" if f32w else
                   "kpipe::kfu_pipe_kernel (fused, software-pipelined Kuf Kfu product)"),
        "bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": (achieved / peak_tflops) if (achieved and peak_tflops) else None,
        "traffic": traffic, "peak_source": peak_src,
        "launches_timed": mv_cnt, "avg_launch_ms": (mv_ms / mv_cnt) if mv_cnt else None,
        "share_of_step": (mv_ms / ms) if ms > 0 else None,
        "sections_ms": {k: round(v[0], 4) for k, v in prof.items()},
        "gentries_per_s": (n_local * M / (mv_ms / mv_cnt * 1e-3) / 1e9) if mv_cnt else None,
    }
    if not f32w and achieved and peak_tflops:
        # what the FP64 pipe actually executes per Gram entry (SASS count of the tile loop): the DMMA
        # distance (4 ceil((D+1)/4) FMA) + sqrt / exp / polynomial / contractions - the algorithmic D + 2 FMA of
        # `achieved` leave the transcendental work out
        slots = 4 * ((D + 4) // 4) + (17.75 if kern.startswith("matern") else 9.75)
        roofline["executed"] = {"fp64_fma_slots_per_entry": slots,
                                "frac_of_peak": achieved / peak_tflops * slots / (D + 2),
                                "note": "FP64-pipe occupancy by executed FMA slots; `frac` counts algorithmic flop only"}

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
            "config": {
                "workload": f"{args.workload}: N={N} (rows sharded over {world} GPU(s), {n_local} on rank 0), M={M}, "
                            f"D={D}, {kern}, B=1 right-hand side; {desc}",
                "operator": "Kuu + jitter I + Kuf Kfu / noise_variance, matrix-free (Kfu never materialised)",
                "step": "one CG iteration: fused Kuf Kfu product + all-reduce + Kuu product + fused vector update",
                "allreduce": (("one-shot kernel over NVLink peer memory (rank-ordered sum)" if ctx.peer_allreduce
                               else "ncclAllReduce") if world > 1 else "none (one rank)"),
                "l2": "inputs larger than L2: prepared X shard %.0f MB + Kuu %.0f MB streamed every iteration (126 MB L2)"
                      % (n_local * (4 * ((D + 4) // 4)) * esize / 1e6, M * M * esize / 1e6),
                "seconds_per_solve": f"{ms_max * 1e-3:.4f} s for {args.steps} iterations (threshold 0, fixed count)",
                "f_alg_per_iteration": f_alg_iteration(N, M, D),
                "fp64_frac_whole_iteration": (f_alg_iteration(N, M, D) * its / world / 1e12 / peak_tflops)
                if peak_tflops else None,
                "hbm_gbs_measured": hbm,
            },
            "roofline": roofline,
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
