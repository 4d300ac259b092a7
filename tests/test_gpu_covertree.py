"""Cover tree on the device (cggp_b200.CoverTree, csrc/covertree.cu) against
  * the golden vectors of the reference's unmodified cggp/covertree.py (tests/golden/covertree_golden.npz), and
  * the oracle restatement (oracle/covertree.py) on larger seeded inputs where many parents run concurrently,
bit for bit: node points of every level, centroids, leaf memberships in order, per-leaf means and counts."""
import os
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "covertree_golden.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files})


def case_kwargs(name):
    return {k.split("/kw_")[1]: GOLD[k].item() for k in GOLD.files if k.startswith(name + "/kw_")}


def build(x, y, **kw):
    import cggp_b200

    dev = torch.device("cuda", 0)
    return cggp_b200.CoverTree(None, (torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev)), **kw)


def check_against(tree, level_sizes, level_points, centroids, means, counts, member_sizes, members):
    sizes = [tree.level_size(lv) for lv in range(tree.num_levels)]
    assert sizes == list(level_sizes)
    pts = torch.cat([tree.level_points(lv) for lv in range(tree.num_levels)]).cpu().numpy()
    first_bad = np.flatnonzero((pts != level_points).any(axis=1))
    assert first_bad.size == 0, f"first differing node {first_bad[:5]} of {len(pts)} (level sizes {sizes})"
    assert np.array_equal(tree.centroids.cpu().numpy(), centroids)
    off, rows = tree.cluster_indices
    assert np.array_equal(np.diff(off.cpu().numpy()), member_sizes)
    assert np.array_equal(rows.cpu().numpy(), members)
    m, c = tree.cluster_mean_and_counts
    assert np.array_equal(c.cpu().numpy(), counts)
    assert np.array_equal(m.cpu().numpy(), means, equal_nan=True)


@pytest.mark.parametrize("cluster", ["0", "1"])   # one CTA per parent / one cluster of 8 CTAs per parent (few parents)
@pytest.mark.parametrize("name", CASES)
def test_covertree_vs_reference_golden(name, cluster, monkeypatch):
    monkeypatch.setenv("CGGP_CT_CLUSTER", cluster)
    x, y = GOLD[f"{name}/x"], GOLD[f"{name}/y"]
    tree = build(x, y, **case_kwargs(name))
    check_against(tree, GOLD[f"{name}/level_sizes"], GOLD[f"{name}/level_points"], GOLD[f"{name}/centroids"],
                  GOLD[f"{name}/means"], GOLD[f"{name}/counts"], GOLD[f"{name}/member_sizes"], GOLD[f"{name}/members"])
    for lv, r in enumerate(GOLD[f"{name}/level_radius"]):
        if tree.level_size(lv):
            assert tree.level_radius(lv) == r


@pytest.mark.parametrize("n,d,res,kw", [
    (40000, 2, 0.12, {}),                      # ~1500 leaves: hundreds of parents per wave at the fine levels
    (20000, 3, 0.35, {}),
    (30000, 2, 0.15, {"voronoi": False}),      # the rows a node took are its data: pool lists, two passes
    (30000, 2, 0.15, {"lloyds": False}),
    (50000, 1, 0.002, {}),                     # D = 1: NumPy's mean over rows is pairwise
])
def test_covertree_vs_oracle_many_parents(n, d, res, kw, monkeypatch):
    from oracle import covertree as oct_

    rng = np.random.default_rng(n + d)
    x = rng.uniform(-3, 3, (n, d))
    y = rng.standard_normal((n, 1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = oct_.CoverTree(None, (x, y), spatial_resolution=res, **kw)
    wm, wc = want.cluster_mean_and_counts
    members = want.cluster_indices
    for cluster in ("0", "1", None):   # None = the default choice per wave
        if cluster is None:
            monkeypatch.delenv("CGGP_CT_CLUSTER", raising=False)
        else:
            monkeypatch.setenv("CGGP_CT_CLUSTER", cluster)
        tree = build(x, y, spatial_resolution=res, **kw)
        check_against(tree, [len(lv) for lv in want.levels],
                      np.concatenate([np.stack([nd.point for nd in lv]) for lv in want.levels if lv]),
                      want.centroids, wm, wc, [len(m) for m in members], np.concatenate(members))


def test_covertree_update_inducing_parameters_and_views():
    import cggp_b200

    name = "uniform2d"
    x, y = GOLD[f"{name}/x"], GOLD[f"{name}/y"]
    dev = torch.device("cuda", 0)
    data = (torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev))
    iv, means, counts = cggp_b200.covertree_update_inducing_parameters(None, data, None, 0.8)
    keep = GOLD[f"{name}/counts"].reshape(-1) != 0
    assert np.array_equal(iv.cpu().numpy(), GOLD[f"{name}/centroids"][keep])
    assert np.array_equal(means.cpu().numpy(), GOLD[f"{name}/means"][keep])
    assert np.array_equal(counts.cpu().numpy(), GOLD[f"{name}/counts"][keep])
    tree = cggp_b200.CoverTree(None, data, spatial_resolution=0.8)
    assert len(tree.nodes) == int(GOLD[f"{name}/level_sizes"].sum())
    leaf = tree.levels[-1][3]
    assert leaf.parent is tree.levels[-2][leaf.parent.index] and leaf.radius == GOLD[f"{name}/level_radius"][-1]
    ys = tree.cluster_ys
    assert [len(v) for v in ys] == GOLD[f"{name}/member_sizes"].tolist()
    with pytest.raises(TypeError):
        cggp_b200.CoverTree(None, (data[0].float(), data[1]), spatial_resolution=0.8)


def test_covertree_large_is_a_partition():
    """Size-independent properties at N = 2M rows (D = 2): every row in exactly one leaf, counts and means agree with
    the memberships, every node lies within its parent's radius."""
    n = 2_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((n, 2), dtype=torch.float64, device="cuda", generator=g) * 20.0 - 10.0
    y = torch.randn((n, 1), dtype=torch.float64, device="cuda", generator=g)
    import cggp_b200

    tree = cggp_b200.CoverTree(None, (x, y), spatial_resolution=0.25)
    off, rows = tree.cluster_indices
    assert rows.numel() == n and torch.equal(torch.sort(rows).values, torch.arange(n, device="cuda"))
    means, counts = tree.cluster_mean_and_counts
    assert counts.sum().item() == n and torch.equal(counts.reshape(-1).long(), off[1:] - off[:-1])
    total = (means.reshape(-1) * counts.reshape(-1))[counts.reshape(-1) > 0].sum().item()
    assert abs(total - y.sum().item()) <= 1e-9 * n
    for lv in range(1, tree.num_levels):
        pts, parents = tree.level_points(lv, with_parents=True)
        up = tree.level_points(lv - 1)[torch.as_tensor(parents, device="cuda")]
        assert ((pts - up).norm(dim=1) <= tree.level_radius(lv - 1) * (1 + 1e-12)).all()
