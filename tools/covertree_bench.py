"""Cover-tree construction time on the device (cggp_b200.CoverTree) for synthetic geospatial-shaped rows, next to the
oracle restatement of the reference on a bounded sample.  Usage: python tools/covertree_bench.py [n] [d] [resolution]"""
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import cggp_b200  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    res = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((n, d), dtype=torch.float64, device="cuda", generator=g) * 20.0 - 10.0
    y = torch.randn((n, 1), dtype=torch.float64, device="cuda", generator=g)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tree = cggp_b200.CoverTree(None, (x, y), spatial_resolution=res)
        means, counts = tree.cluster_mean_and_counts
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sizes = [tree.level_size(lv) for lv in range(tree.num_levels)]
        print(f"device: n={n} D={d} resolution={res}: {t1 - t0:.3f} s, levels {sizes}, "
              f"empty leaves {int((counts == 0).sum())}", flush=True)
    ns = min(n, 60000)
    from oracle import covertree as oct_

    xs, ys = x[:ns].cpu().numpy(), y[:ns].cpu().numpy()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = oct_.CoverTree(None, (xs, ys), spatial_resolution=res)
    t1 = time.perf_counter()
    t2 = time.perf_counter()
    tree = cggp_b200.CoverTree(None, (x[:ns], y[:ns]), spatial_resolution=res)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    same = np.array_equal(tree.centroids.cpu().numpy(), want.centroids)
    print(f"sample n={ns}: NumPy restatement of the reference {t1 - t0:.2f} s, device {t3 - t2:.3f} s, "
          f"leaves {len(want.levels[-1])}, identical centroids: {same}")


if __name__ == "__main__":
    main()
