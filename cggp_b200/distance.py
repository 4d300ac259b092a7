"""Mirror of ``cggp/distance.py``: point-to-point distances used for cluster assignment.

``euclid_distance((x, y))`` (distance.py:9-11) and ``create_distance_fn(kernel, distance_type)`` with
``"euclidean" | "covariance" | "correlation"`` (distance.py:14-34) keep the reference call convention
``distance_fn((centroids [M, D], point [D])) -> [M]`` (selection.py:24-31) and also accept a batch of points
``[N, D]`` (-> ``[M, N]``).  The returned callables carry ``distance_type`` / ``kernel`` so that
``selection.kmeans_indices_and_distances`` can run the fused N x M distance + argmin kernel instead of a
per-point map.
"""
from __future__ import annotations

from typing import Literal

import torch

from . import _lib
from .kernels import Stationary, kernel_matrix, prepare_points

DistanceType = Literal["euclidean", "covariance", "correlation"]


def _pairwise(kernel, distance_type, x, y):
    x = _lib.as_device_tensor(x)
    y = _lib.as_device_tensor(y, x.dtype)
    squeeze_x = x.dim() == 1
    squeeze_y = y.dim() == 1
    x2 = x.reshape(1, -1) if squeeze_x else x
    y2 = y.reshape(1, -1) if squeeze_y else y
    code = _lib.DISTANCE_CODES[distance_type]
    if distance_type in ("euclidean", "sqeuclidean"):
        A, B = prepare_points(x2, 1.0), prepare_points(y2, 1.0)
        out = kernel_matrix(_lib.SE, 1.0, A, B, output=_lib.OUT_DISTANCE, distance=code)
    else:
        A, B = kernel.prepare(x2), kernel.prepare(y2, x2.dtype)
        out = kernel_matrix(kernel.kind, kernel.variance, A, B, output=_lib.OUT_DISTANCE, distance=code)
    if squeeze_y:
        out = out[:, 0]
    if squeeze_x:
        out = out[0]
    return out


def euclid_distance(args):
    """``||x - y||_2`` over the last axis, difference form (distance.py:9-11)."""
    x, y = args
    return _pairwise(None, "euclidean", x, y)


euclid_distance.distance_type = "euclidean"
euclid_distance.kernel = None


def create_distance_fn(kernel: Stationary, distance_type: DistanceType):
    def cov(args):  # distance.py:15-22: k(x,x) + k(y,y) - 2 k(x,y)
        x, y = args
        return _pairwise(kernel, "covariance", x, y)

    def cor(args):  # distance.py:24-30: 1 - k(x,y) / sqrt(k(x,x) k(y,y))
        x, y = args
        return _pairwise(kernel, "correlation", x, y)

    cov.distance_type, cov.kernel = "covariance", kernel
    cor.distance_type, cor.kernel = "correlation", kernel
    functions = {"covariance": cov, "correlation": cor, "euclidean": euclid_distance}
    return functions[distance_type]
