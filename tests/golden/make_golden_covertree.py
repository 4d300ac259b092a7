"""Golden vectors of the cover tree: run the reference's UNMODIFIED ``cggp/covertree.py`` (from /root/reference,
read-only; its ``tensorflow`` / ``matplotlib`` imports are satisfied by the stand-ins in ``tests/golden/_shim``, it
uses neither) on seeded inputs and store inputs + outputs in ``tests/golden/covertree_golden.npz`` (committed).

Run in the build container only:  ``python tests/golden/make_golden_covertree.py``.
Every case is run twice with the same rows: once with the real targets (centroids, per-leaf means and counts) and once
with the row number as the target, which makes the reference reveal WHICH rows every leaf holds, in its order.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/cggp"
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, REF)

import covertree as ref_ct  # noqa: E402  (reference, unmodified)

assert ref_ct.__file__.startswith(REF)


def cases():
    out = {}
    rng = np.random.default_rng(20260)
    out["uniform2d"] = (rng.uniform(-5, 5, (3000, 2)), dict(spatial_resolution=0.8))
    out["gauss3d"] = (rng.standard_normal((2500, 3)), dict(spatial_resolution=0.7))
    out["line1d"] = (rng.uniform(0, 1, (1200, 1)), dict(spatial_resolution=0.05))
    out["gauss11d"] = (rng.standard_normal((1500, 11)), dict(spatial_resolution=2.0))
    out["levels_no_resolution"] = (rng.standard_normal((1000, 2)), dict(num_levels=4))
    out["no_lloyds"] = (rng.uniform(-1, 1, (1500, 2)), dict(spatial_resolution=0.2, lloyds=False))
    out["no_voronoi"] = (rng.uniform(-1, 1, (1500, 2)), dict(spatial_resolution=0.2, voronoi=False))
    out["neither"] = (rng.standard_normal((900, 3)), dict(spatial_resolution=0.9, lloyds=False, voronoi=False))
    grid = rng.integers(0, 12, (800, 2)).astype(np.float64) * 0.25      # exact duplicates and equidistant rows
    out["duplicates"] = (grid, dict(spatial_resolution=0.3))
    out["single_row"] = (rng.standard_normal((1, 2)), dict(num_levels=3))
    out["five_rows"] = (rng.standard_normal((5, 3)), dict(spatial_resolution=0.5))
    out["root_only"] = (rng.standard_normal((300, 2)), dict(num_levels=1))
    out["wide40d"] = (rng.standard_normal((400, 40)) * 0.2, dict(spatial_resolution=0.9))
    out["wide9d"] = (rng.uniform(-1, 1, (700, 9)), dict(spatial_resolution=1.0))
    return out, rng


def main():
    store = {}
    all_cases, rng = cases()
    for name, (x, kw) in all_cases.items():
        n = x.shape[0]
        y = rng.standard_normal((n, 1))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tree = ref_ct.CoverTree(None, (x.copy(), y.copy()), **kw)
            means, counts = tree.cluster_mean_and_counts
            rows = np.arange(n, dtype=np.float64)[:, None]
            tree_rows = ref_ct.CoverTree(None, (x.copy(), rows), **kw)
        assert np.array_equal(tree.centroids, tree_rows.centroids)
        members = [node.data[1][:, 0].astype(np.int64) for node in tree_rows.levels[-1]]
        store[f"{name}/x"] = x
        store[f"{name}/y"] = y
        for k, v in kw.items():
            store[f"{name}/kw_{k}"] = np.asarray(v)
        store[f"{name}/centroids"] = tree.centroids
        store[f"{name}/means"] = means
        store[f"{name}/counts"] = counts
        store[f"{name}/level_sizes"] = np.array([len(lv) for lv in tree.levels], dtype=np.int64)
        store[f"{name}/level_points"] = np.concatenate([np.stack([nd.point for nd in lv]) for lv in tree.levels if lv])
        store[f"{name}/level_radius"] = np.array([lv[0].radius if lv else np.nan for lv in tree.levels])
        store[f"{name}/member_sizes"] = np.array([len(m) for m in members], dtype=np.int64)
        store[f"{name}/members"] = np.concatenate(members) if members else np.zeros(0, np.int64)
        print(f"{name}: n={n} D={x.shape[1]} levels={[len(lv) for lv in tree.levels]} "
              f"empty leaves={int((counts == 0).sum())}")
    np.savez_compressed(os.path.join(HERE, "covertree_golden.npz"), **store)


if __name__ == "__main__":
    main()
