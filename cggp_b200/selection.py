"""Nearest-centre assignment and cluster statistics: the step that feeds ``(Z, pseudo_u, cluster_counts)`` to the
models (``cggp/selection.py:14-32``, ``cggp/optimize.py:41-98``), as one fused N x M distance + argmin kernel
(``cggp_nearest_center``) and an atomic segmented sum (``cggp_cluster_stats``)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

from . import _lib
from .distance import euclid_distance
from .kernels import prepare_points


def _nearest(points, centroids, distance_type, kernel):
    points = _lib.as_device_tensor(points)
    centroids = _lib.as_device_tensor(centroids, points.dtype)
    if distance_type in ("euclidean", "sqeuclidean"):
        PX, PZ = prepare_points(points, 1.0), prepare_points(centroids, 1.0)
        kind, variance = _lib.SE, 1.0
    else:
        PX, PZ = kernel.prepare(points), kernel.prepare(centroids, points.dtype)
        kind, variance = kernel.kind, kernel.variance
    ctx = _lib.context(points.device)
    ctx.use_current_stream()
    idx = torch.empty((PX.n,), dtype=torch.int64, device=points.device)
    dist = torch.empty((PX.n,), dtype=points.dtype, device=points.device)
    ctx.check(ctx.lib.cggp_nearest_center(
        ctx.handle, _lib.dtype_code(points.dtype), kind, variance, _lib.DISTANCE_CODES[distance_type],
        _lib.ptr(PX.P), _lib.ptr(PX.norms), PX.n, _lib.ptr(PZ.P), _lib.ptr(PZ.norms), PZ.n, PX.D, PX.ldp,
        _lib.ptr(idx), _lib.ptr(dist)))
    return idx, dist


def kmeans_indices_and_distances(centroids, points, distance_fn: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """selection.py:14-32: per point the index of the nearest centroid (first minimum) and that distance."""
    if distance_fn is None:
        distance_fn = euclid_distance
    dtype = getattr(distance_fn, "distance_type", None)
    if dtype is None:
        raise TypeError("distance_fn must come from cggp_b200.distance (euclid_distance / create_distance_fn)")
    return _nearest(points, centroids, dtype, getattr(distance_fn, "kernel", None))


def cluster_stats(indices, y, m) -> Tuple[torch.Tensor, torch.Tensor]:
    """counts[j] and sum of y over cluster j (float tensors [m])."""
    y = _lib.as_device_tensor(y).reshape(-1).contiguous()
    indices = _lib.as_device_tensor(indices).to(torch.int64).contiguous()
    ctx = _lib.context(y.device)
    ctx.use_current_stream()
    counts = torch.empty((m,), dtype=y.dtype, device=y.device)
    sums = torch.empty((m,), dtype=y.dtype, device=y.device)
    ctx.check(ctx.lib.cggp_cluster_stats(ctx.handle, _lib.dtype_code(y.dtype), _lib.ptr(indices), _lib.ptr(y),
                                         y.numel(), m, _lib.ptr(counts), _lib.ptr(sums)))
    return counts, sums


def kmeans_update_inducing_parameters(model, data, distance_fn: Optional[Callable], clustering_fn: Callable):
    """optimize.py:81-98: ``(new_iv, u = cluster means of y, counts)`` (empty clusters give 0/0 = nan, as there)."""
    x, y = data
    new_iv = _lib.as_device_tensor(clustering_fn())
    m = new_iv.shape[0]
    indices, _ = kmeans_indices_and_distances(new_iv, x, distance_fn=distance_fn)
    counts, sums = cluster_stats(indices, y, m)
    u = (sums / counts)[:, None]
    return new_iv, u, counts[:, None]


def nearest_center_update(iv, data):
    """optimize.py:50-78 (OIPS / uniform / greedy): assignment by ``argmin(square_distance(iv, inputs), axis=0)``,
    per-cluster mean of the outputs and counts with empty clusters replaced by 1."""
    x, y = data
    iv = _lib.as_device_tensor(iv)
    idx, _ = _nearest(x, iv, "sqeuclidean", None)
    counts, sums = cluster_stats(idx, y, iv.shape[0])
    means = sums / counts
    new_counts = torch.where(counts != 0, counts, torch.ones_like(counts))
    return iv, means, new_counts
