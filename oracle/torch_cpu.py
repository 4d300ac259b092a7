"""Multi-threaded CPU port of the matrix-free CG iteration (torch CPU ops, float64) used ONLY as the timed CPU
baseline of bench.py (`cpu_baseline`, `--impl reference`).  TEST / BENCH INFRASTRUCTURE, not product code.

Same arithmetic as oracle/gpflow_restated.py + oracle/models.py::sgpr_operator + oracle/cg.py (the restated reference
path: GPflow expanded squared distance, Matern/SE K_r2, `state.p @ A`, the cg_step of cggp/conjugate_gradient.py:64-85),
expressed with torch CPU tensors so that the element-wise kernel evaluation uses every host core
(NumPy's ufuncs are single-threaded); checked against the NumPy oracle in tests/test_oracle_gpflow.py.
The reference itself (TensorFlow + GPflow) is not installable in this image, hence kind = "port".
"""
from __future__ import annotations

import math

import torch


def kernel_matrix(name, variance, ls, X, Z):
    Xs, Zs = X / ls, Z / ls
    r2 = -2.0 * (Xs @ Zs.T) + ((Xs * Xs).sum(-1)[:, None] + (Zs * Zs).sum(-1)[None, :])
    if name == "se":
        return variance * torch.exp(-0.5 * r2)
    r = torch.sqrt(torch.clamp_min(r2, 1e-36))
    if name == "matern12":
        return variance * torch.exp(-r)
    if name == "matern32":
        s = math.sqrt(3.0) * r
        return variance * (1.0 + s) * torch.exp(-s)
    s = math.sqrt(5.0) * r
    return variance * (1.0 + s + (5.0 / 3.0) * r * r) * torch.exp(-s)


def sgpr_operator(name, variance, ls, X, Z, noise, jitter=1e-6, chunk=16384):
    kuu = kernel_matrix(name, variance, ls, Z, Z) + jitter * torch.eye(Z.shape[0], dtype=Z.dtype)

    def matmul(V):
        out = V @ kuu
        acc = torch.zeros_like(V)
        for s in range(0, X.shape[0], chunk):
            K = kernel_matrix(name, variance, ls, X[s:s + chunk], Z)
            acc += (V @ K.T) @ K
        return out + acc / noise

    return matmul


def cg_iterations(matmul, rhs, iters):
    """`iters` steps of cggp/conjugate_gradient.py:64-85 (Eye preconditioner, no refresh) from v = 0."""
    v = torch.zeros_like(rhs)
    r = rhs.clone()
    rz = (r * r).sum(-1, keepdim=True)
    p = r.clone()
    hist = [0.5 * rz[:, 0].clone()]
    for _ in range(iters):
        pA = matmul(p)
        denom = (p * pA).sum(-1, keepdim=True)
        gamma = torch.where(denom <= 1e-16, torch.zeros_like(denom), rz / denom)
        v = v + gamma * p
        r = r - gamma * pA
        new_rz = (r * r).sum(-1, keepdim=True)
        upd = torch.where(rz <= 1e-16, torch.zeros_like(p), p * new_rz / rz)
        p = r + upd
        rz = new_rz
        hist.append(0.5 * rz[:, 0].clone())
    return v, torch.stack(hist)
