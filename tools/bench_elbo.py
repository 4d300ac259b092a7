"""Seconds per SGPR.elbo (GPflow's collapsed bound, the objective behind cli_utils.sgpr_class) and per CDGP elbo at the c3
shape on one GPU (development aid): where the N M^2 Gram runs on the library's DMMA SYRK (cggp_kuf_gram)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cggp_b200 as cb
from cggp_b200 import selection


def main():
    N, M, D = 2_000_000, 4096, 11
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(N, D, dtype=torch.float64, device="cuda", generator=g)
    y = torch.sin(X.sum(-1, keepdim=True)) + 0.3 * torch.randn(N, 1, dtype=torch.float64, device="cuda", generator=g)
    Z = X[torch.randperm(N, device="cuda", generator=g)[:M]].clone()
    k = cb.Matern52(1.0, [1.0] * D)
    with torch.no_grad():
        model = cb.sgpr_class((X, y), k, cb.Gaussian(0.1), Z)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e = model.elbo()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"SGPR.elbo at c3 (N={N}, M={M}, D={D}): {dt:.3f} s, value {float(e):.6e}; "
                  f"{2.0 * N * M * M / dt / 1e12:.1f} TFLOP/s in N M^2 terms end to end", flush=True)
        # CDGP objective on a minibatch of 5000 (the reference's training step, optimize.py:198-254; batch 1000-5000)
        _, means, counts = selection.nearest_center_update(Z, (X, y))
        cd = cb.cdgp_class(k, cb.Gaussian(0.1), Z, error_threshold=1e-6, cluster_counts=counts[:, None],
                           pseudo_u=torch.nan_to_num(means)[:, None], num_data=N, num_probes=5)
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e = cd.elbo((X[:5000], y[:5000]))
            torch.cuda.synchronize()
            print(f"CDGP elbo, minibatch 5000, M={M}, 5 probes: {time.perf_counter() - t0:.3f} s, value {float(e):.6e}",
                  flush=True)


if __name__ == "__main__":
    main()
