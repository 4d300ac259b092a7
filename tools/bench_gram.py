"""Timing of cggp_kuf_gram (Kuf Kfu [M, M] by row chunks + DMMA SYRK): python tools/bench_gram.py [M] [rows] [D]
(chunk rows: env CGGP_GRAM_ROWS, read once per process)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cggp_b200 as cb
from cggp_b200.kernels import kuf_gram


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.rand(rows, D, dtype=torch.float64, device="cuda", generator=g) * M ** 0.5
    Z = torch.rand(M, D, dtype=torch.float64, device="cuda", generator=g) * M ** 0.5
    k = cb.Matern32(1.0, [1.0] * D)
    PZ, PX = k.prepare(Z), k.prepare(X)
    G = kuf_gram(k.kind, k.variance, PZ, PX)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    G = kuf_gram(k.kind, k.variance, PZ, PX)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    full = 2.0 * rows * M * M
    print(f"gram M={M} rows={rows} D={D} chunk_rows={os.environ.get('CGGP_GRAM_ROWS', 'default')}: {ms:.2f} ms, "
          f"{full / ms / 1e9:.2f} TFLOP/s in N M^2 terms, {full / 2 / ms / 1e9:.2f} executed (lower tiles); "
          f"sym err {float((G - G.t()).abs().max()):.1e}", flush=True)


if __name__ == "__main__":
    main()
