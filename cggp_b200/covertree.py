"""Cover-tree inducing-point selection (``cggp/covertree.py:25-176``) on the device.

``CoverTree`` keeps the reference's constructor arguments and result properties (``centroids``, ``cluster_ys``,
``cluster_mean_and_counts``, ``levels``, ``nodes``); the rows stay in HBM and the construction runs in
``cggp_covertree_build`` (csrc/covertree.cu): the reference's greedy order of operations and NumPy's orders of summation
are reproduced, so centroids, counts and means are the reference's bit for bit (tests/test_gpu_covertree.py checks
that against golden vectors of the unmodified reference file).  float64 rows (the reference's default precision).
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Callable, List, Optional, Tuple

import torch

from . import _lib


class CoverTreeNode:
    """Read-only view of one node (the reference's ``CoverTreeNode`` without the copies of the rows)."""

    __slots__ = ("point", "radius", "parent", "level", "index")

    def __init__(self, point, radius, parent, level, index):
        self.point, self.radius, self.parent, self.level, self.index = point, radius, parent, level, index


class CoverTree:
    def __init__(self, distance: Optional[Callable], data, spatial_resolution: Optional[float] = None,
                 num_levels: Optional[int] = 1, lloyds: bool = True, voronoi: bool = True, plotting: bool = False):
        if distance is not None:
            # covertree.py:36-38: the reference ignores it too and says so
            warnings.warn("Distance function will be ignored and instead the Euclidean norm will be used.")
        if plotting:
            raise NotImplementedError("plotting copies of the rows are not kept on the device")
        x, y = data
        x = _lib.row_major(_lib.as_device_tensor(x))
        if x.dtype != torch.float64:
            raise TypeError("cggp_b200.CoverTree takes float64 rows (the reference's default_float)")
        y = _lib.as_device_tensor(y, device=x.device)
        if y.dim() == 1:
            y = y[:, None]
        if y.shape[0] != x.shape[0]:
            raise ValueError("inputs and targets differ in their number of rows")
        self.x, self.y = x, y
        self._ctx = _lib.context(x.device)
        self._ctx.use_current_stream()
        handle = C.c_void_p()
        self._ctx.check(self._ctx.lib.cggp_covertree_build(
            self._ctx.handle, _lib.F64, _lib.ptr(x), x.shape[0], x.shape[1], x.stride(0),
            float(spatial_resolution) if spatial_resolution is not None else 0.0,
            int(num_levels) if num_levels is not None else 1, int(bool(lloyds)), int(bool(voronoi)), C.byref(handle)))
        self._handle = handle
        self.num_levels = int(self._ctx.lib.cggp_covertree_num_levels(handle))
        self._levels = None
        self._members = None

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._ctx.lib.cggp_covertree_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ---- per-level queries ----
    def level_size(self, level: int) -> int:
        return int(self._ctx.lib.cggp_covertree_level_size(self._handle, level % self.num_levels))

    def level_radius(self, level: int) -> float:
        r = C.c_double(0.0)
        self._ctx.check(self._ctx.lib.cggp_covertree_level_radius(self._handle, level % self.num_levels, C.byref(r)))
        return r.value

    def level_points(self, level: int, with_parents: bool = False):
        """``[node.point for node in levels[level]]`` as a device tensor ``[size, D]`` (and the parent indices)."""
        level %= self.num_levels
        m = self.level_size(level)
        out = torch.empty((m, self.x.shape[1]), dtype=torch.float64, device=self.x.device)
        parents = (C.c_int32 * max(m, 1))()
        self._ctx.use_current_stream()
        self._ctx.check(self._ctx.lib.cggp_covertree_level_points(self._ctx.handle, self._handle, level, _lib.ptr(out),
                                                                  out.stride(0) if m else self.x.shape[1], parents))
        return (out, list(parents[:m])) if with_parents else out

    # ---- the reference's properties ----
    @property
    def centroids(self) -> torch.Tensor:
        """covertree.py:159-161."""
        return self.level_points(-1)

    @property
    def cluster_indices(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Row numbers of every leaf in the reference's order: ``(offsets [M + 1], rows [N])`` (int64, device)."""
        if self._members is None:
            m = self.level_size(-1)
            off = torch.empty((m + 1,), dtype=torch.int64, device=self.x.device)
            rows = torch.empty((self.x.shape[0],), dtype=torch.int64, device=self.x.device)
            self._ctx.use_current_stream()
            self._ctx.check(self._ctx.lib.cggp_covertree_leaf_members(self._ctx.handle, self._handle, _lib.ptr(off),
                                                                      _lib.ptr(rows)))
            self._members = (off, rows)
        return self._members

    @property
    def cluster_ys(self) -> List[torch.Tensor]:
        """covertree.py:163-166: the targets of every leaf."""
        off, rows = self.cluster_indices
        ys = self.y[rows]
        bounds = off.tolist()
        return [ys[bounds[k]:bounds[k + 1]] for k in range(len(bounds) - 1)]

    @property
    def cluster_mean_and_counts(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """covertree.py:168-176: ``(means [M, 1], counts [M, 1])`` in the dtype of the targets."""
        if self.y.shape[1] != 1 or self.y.dtype != torch.float64:
            raise TypeError("cluster_mean_and_counts takes float64 targets of shape [N, 1]")
        m = self.level_size(-1)
        means = torch.empty((m, 1), dtype=torch.float64, device=self.x.device)
        counts = torch.empty((m, 1), dtype=torch.float64, device=self.x.device)
        self._ctx.use_current_stream()
        self._ctx.check(self._ctx.lib.cggp_covertree_cluster_stats(
            self._ctx.handle, self._handle, _lib.F64, _lib.ptr(self.y), self.y.stride(0), _lib.ptr(means),
            _lib.ptr(counts)))
        return means, counts

    @property
    def levels(self) -> List[List[CoverTreeNode]]:
        if self._levels is None:
            self._levels = []
            for lv in range(self.num_levels):
                pts, parents = self.level_points(lv, with_parents=True)
                radius = self.level_radius(lv)
                prev = self._levels[lv - 1] if lv else None
                self._levels.append([CoverTreeNode(pts[k], radius, prev[parents[k]] if prev else None, lv, k)
                                     for k in range(pts.shape[0])])
        return self._levels

    @property
    def nodes(self) -> List[CoverTreeNode]:
        return [node for level in self.levels for node in level]


def covertree_update_inducing_parameters(model, data, distance_fn, spatial_resolution: float):
    """cggp/optimize.py:19-39: ``(new_iv, means, counts)`` of the leaves, empty leaves dropped."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = CoverTree(distance_fn, data, spatial_resolution=spatial_resolution)
    new_iv = tree.centroids
    means, counts = tree.cluster_mean_and_counts
    keep = (counts != 0.0).reshape(-1)
    return new_iv[keep], means[keep], counts[keep]
