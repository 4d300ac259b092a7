"""The cover-tree restatement (oracle/covertree.py) against the golden vectors produced by the reference's unmodified
``cggp/covertree.py`` (tests/golden/make_golden_covertree.py), bit for bit; and the spelled-out NumPy summation orders
the CUDA engine implements against NumPy itself."""
import os

import numpy as np
import pytest

from oracle import covertree as oct_

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "covertree_golden.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files})


def case_kwargs(name):
    kw = {}
    for k in GOLD.files:
        if k.startswith(name + "/kw_"):
            v = GOLD[k]
            kw[k.split("/kw_")[1]] = v.item()
    return kw


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_golden(name):
    x, y = GOLD[f"{name}/x"], GOLD[f"{name}/y"]
    tree = oct_.CoverTree(None, (x, y), **case_kwargs(name))
    assert [len(lv) for lv in tree.levels] == GOLD[f"{name}/level_sizes"].tolist()
    pts = np.concatenate([np.stack([nd.point for nd in lv]) for lv in tree.levels if lv])
    assert np.array_equal(pts, GOLD[f"{name}/level_points"])
    assert np.array_equal(tree.centroids, GOLD[f"{name}/centroids"])
    means, counts = tree.cluster_mean_and_counts
    assert np.array_equal(counts, GOLD[f"{name}/counts"])
    assert np.array_equal(means, GOLD[f"{name}/means"], equal_nan=True)
    members = tree.cluster_indices
    assert [len(m) for m in members] == GOLD[f"{name}/member_sizes"].tolist()
    assert np.array_equal(np.concatenate(members), GOLD[f"{name}/members"])
    for lv, r in zip(tree.levels, GOLD[f"{name}/level_radius"]):
        if lv:
            assert lv[0].radius == r


def test_numpy_summation_orders_as_spelled_out():
    rng = np.random.default_rng(7)
    for d in (1, 2, 3, 7, 8, 9, 11, 17, 40, 130, 300):
        rows = rng.standard_normal((64, d)) * 10.0 ** rng.integers(-4, 4, size=(64, d))
        p = rng.standard_normal(d)
        diff = p - rows
        want = np.linalg.norm(diff, axis=-1)
        got = np.array([np.sqrt(oct_.pairwise_sum(r * r)) for r in diff])
        assert np.array_equal(want, got), d
    for n, d in ((1, 3), (7, 2), (1000, 11), (5000, 1), (129, 1), (300, 40)):
        rows = rng.standard_normal((n, d)) * 10.0 ** rng.integers(-3, 3, size=(n, 1))
        assert np.array_equal(rows.mean(axis=-2), oct_.ordered_row_mean(rows)), (n, d)
    for n in (1, 7, 8, 9, 128, 129, 1000, 4097):
        y = rng.standard_normal((n, 1)) * 10.0 ** rng.integers(-3, 3, size=(n, 1))
        assert np.mean(y) == oct_.pairwise_sum(y[:, 0]) / n


def test_one_dimensional_norm_is_an_fma_chain_here():
    rng = np.random.default_rng(8)
    for d in (1, 2, 3, 11, 15):
        for _ in range(50):
            v = rng.standard_normal(d) * 10.0 ** rng.integers(-4, 4, size=d)
            assert np.linalg.norm(v) == oct_.fma_chain_norm(v)
