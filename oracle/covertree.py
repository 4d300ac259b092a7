"""TEST / BENCH INFRASTRUCTURE ONLY - never imported by the product (cggp_b200/).

CPU restatement (NumPy) of the reference's cover-tree inducing-point selection, ``cggp/covertree.py:25-176``
(SURVEY.md section 8(f), rank 4).  Pinned: ``tests/golden/covertree_golden.npz`` holds inputs and outputs of the
UNMODIFIED reference file (run over the NumPy shims by ``tests/golden/make_golden_covertree.py``);
``tests/test_oracle_covertree.py`` checks this restatement against them bit for bit.

The restatement keeps, per node, an ordered list of ROW INDICES into the data instead of the reference's copies of
the rows themselves; everything else follows the reference statement by statement:

  root            covertree.py:49-63   mean of all rows, largest distance to it, optional re-scaling of the radius
  greedy pass     covertree.py:66-101  per parent, while it has rows left: first remaining row -> (Lloyd) mean of the
                                       parent's rows within `radius` of it, rejected for the row itself when it comes
                                       within `radius` of an existing child of a neighbouring parent -> the new child
                                       takes (removes) every row within `radius` from all neighbouring parents
  neighbours      covertree.py:102-116 children of the parent's neighbours within neighbor_factor[level] * radius
  Voronoi pass    covertree.py:117-155 every row goes to the nearest candidate child (first minimum), parents in order
  outputs         covertree.py:159-176 centroids, cluster_ys, cluster_mean_and_counts

Arithmetic that decides results (NumPy semantics, which the CUDA engine reproduces):
  * row distance   sqrt(pairwise_sum_d (x_d - p_d)^2): ``np.linalg.norm(x - y, axis=-1)`` = add.reduce over the
                   contiguous last axis = NumPy's pairwise summation (< 8 terms: left to right; <= 128: 8 accumulators)
  * 1-D norm       ``np.linalg.norm(point - child.point)`` of covertree.py:79 goes through ``dot``: for D <= 15 a
                   left-to-right chain of fused multiply-adds on the build container's NumPy/OpenBLAS (a SIMD order
                   the engine does not chase for wider rows; the value is only compared with `<`, so a last-bit
                   difference matters for exact ties alone)
  * mean of rows   ``x.mean(axis=-2)``: rows added one after the other (D >= 2); for D = 1 the axis is contiguous and
                   the sum is pairwise; then one division by the count
  * mean of y      ``np.mean(y)`` over a contiguous [n, 1] array: pairwise summation / n
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np


def row_distance(point, rows):
    """covertree.py:40-43 ``distance_fn`` for one point against [n, D] rows."""
    return np.linalg.norm(point - rows, axis=-1)


class Node:
    def __init__(self, point, radius, parent, idx, r_neighbors=None):
        self.point = point
        self.radius = radius
        self.parent = parent
        self.idx = idx                      # ordered row indices (the reference's node.data)
        self.children: List["Node"] = []
        self.r_neighbors = [self] if r_neighbors is None else r_neighbors
        self.vor: Optional[np.ndarray] = None   # the reference's node.voronoi_data


class CoverTree:
    """Same constructor arguments and properties as the reference class (``distance`` is ignored there too)."""

    def __init__(self, distance, data, spatial_resolution=None, num_levels=1, lloyds=True, voronoi=True,
                 plotting=False):
        x, y = data
        x = np.asarray(x)
        y = np.asarray(y)
        self.x, self.y = x, y
        n = x.shape[0]
        root_mean = x.mean(axis=-2)                                         # :49
        max_radius = np.max(row_distance(root_mean, x))                     # :50-51
        if spatial_resolution is not None:                                  # :53-55
            num_levels = math.ceil(math.log2(max_radius / spatial_resolution)) + 1
            max_radius = spatial_resolution * (2 ** (num_levels - 1))
        root = Node(root_mean, max_radius, None, np.arange(n, dtype=np.int64))
        if voronoi:
            root.vor = root.idx.copy()
        self.levels: List[List[Node]] = [[] for _ in range(num_levels)]
        self.levels[0].append(root)
        neighbor_factor = 4 * (1 - 1 / 2 ** np.arange(num_levels, -1, -1))  # :64

        for level in range(1, num_levels):
            radius = max_radius / (2 ** level)
            for parent in self.levels[level - 1]:                           # greedy pass, :68-101
                while len(parent.idx) > 0:
                    initial_point = x[parent.idx[0]]
                    if lloyds:
                        rows = x[parent.idx]
                        near = row_distance(initial_point, rows) <= radius
                        point = rows[near, :].mean(axis=-2)
                        for r_neighbor in parent.r_neighbors:
                            if any(np.linalg.norm(point - c.point) < radius for c in r_neighbor.children):
                                point = initial_point
                                break
                    else:
                        point = initial_point
                    taken = []
                    for r_neighbor in parent.r_neighbors:
                        inside = row_distance(point, x[r_neighbor.idx]) <= radius
                        taken.append(r_neighbor.idx[inside])
                        r_neighbor.idx = r_neighbor.idx[~inside]
                    child = Node(point, radius, parent, np.concatenate(taken))
                    self.levels[level].append(child)
                    parent.children.append(child)
            for parent in self.levels[level - 1]:                           # neighbours, :102-116
                candidates = [c for r in parent.r_neighbors for c in r.children]
                for child in parent.children:
                    child.r_neighbors = [c for c in candidates
                                         if np.linalg.norm(c.point - child.point, axis=-1)
                                         <= neighbor_factor[level] * radius]
            if voronoi:                                                     # :117-155
                for parent in self.levels[level - 1]:
                    if parent.vor.size == 0:
                        continue
                    candidates = [c for r in parent.r_neighbors for c in r.children]
                    pts = np.stack([c.point for c in candidates])
                    dist = np.linalg.norm(pts[:, None, :] - x[parent.vor][None, :, :], axis=-1)
                    nearest = np.argmin(dist, axis=0)
                    for k, c in enumerate(candidates):
                        if c.vor is None:
                            c.vor = np.empty((0,), dtype=np.int64)
                        c.vor = np.concatenate((c.vor, parent.vor[nearest == k]))
                        c.idx = c.vor.copy()
        self.nodes = [node for lvl in self.levels for node in lvl]

    @property
    def centroids(self):
        return np.stack([node.point for node in self.levels[-1]])

    @property
    def cluster_indices(self):
        """Row indices of every leaf, in order (what the reference carries as copies of the rows)."""
        return [node.idx for node in self.levels[-1]]

    @property
    def cluster_ys(self):
        return [self.y[node.idx] for node in self.levels[-1]]

    @property
    def cluster_mean_and_counts(self):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")   # empty leaf: mean of nothing = nan, as in the reference
            means = [np.mean(self.y[node.idx]) for node in self.levels[-1]]
        counts = [len(node.idx) for node in self.levels[-1]]
        dtype = self.y.dtype
        return np.array(means, dtype=dtype)[..., None], np.array(counts, dtype=dtype)[..., None]


# ---- NumPy's summation orders spelled out (what the CUDA engine implements; checked against NumPy in the tests) ----

def pairwise_sum(a):
    """NumPy's pairwise summation of a contiguous 1-D run (numpy/core/src/umath/loops_utils.h, PW_BLOCKSIZE 128)."""
    n = len(a)
    if n < 8:
        r = a.dtype.type(0.0)
        for v in a:
            r = r + v
        return r
    if n <= 128:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def ordered_row_mean(rows):
    """``rows.mean(axis=-2)`` of a C-contiguous [n, D] array: row after row for D >= 2, pairwise for D = 1."""
    n, d = rows.shape
    if d == 1:
        return np.array([pairwise_sum(rows[:, 0])]) / n
    s = rows[0].copy()
    for i in range(1, n):
        s = s + rows[i]
    return s / n


def fma_chain_norm(v):
    """``np.linalg.norm`` of a 1-D float64 vector (D <= 15) on the build container: sqrt of s = fma(v_i, v_i, s)."""
    import ctypes

    libm = ctypes.CDLL("libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    s = 0.0
    for t in v:
        s = libm.fma(float(t), float(t), s)
    return math.sqrt(s)
