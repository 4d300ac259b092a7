"""NumPy restatement of ``cggp/conjugate_gradient.py``.  TEST INFRASTRUCTURE, not product code.

Pinned against the reference's own unmodified file run over a NumPy shim (``tests/golden/make_golden.py``).
Row convention as in the reference: ``rhs`` is ``[m, n]`` (m independent right-hand sides as ROWS), the product is
``p @ A``.  ``matrix`` may be an ``[n, n]`` array or a callable ``V -> V @ A`` (matrix-free operator), the extension
the new build adds at the same argument position.
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import numpy as np


class CGState(NamedTuple):  # conjugate_gradient.py:10-21
    i: int
    v: np.ndarray
    r: np.ndarray
    p: np.ndarray
    rz: np.ndarray


class EyePreconditioner:
    """conjugate_gradient.py:131-134: ``z = r``, ``rz = sum r^2`` row-wise (keepdims)."""

    def __call__(self, vec, mat):
        return vec, np.sum(np.square(vec), axis=-1, keepdims=True)


class BlockPreconditioner:
    """Block-Jacobi as conjugate_gradient.py:137-157 evidently intends.

    The reference gathers rows of the RHS batch instead of vector elements (``tf.gather(vec, indices)`` with
    ``vec`` of shape [m, n], line 144) and is never instantiated; this restates the intent:
    ``z[:, blk] = A[blk, blk]^-1 r[:, blk]`` via Cholesky (lines 152-154), ``rz = sum z*r`` (line 157).
    ``block_indices``: int array [num_blocks, block_size] partitioning ``range(n)``.
    """

    def __init__(self, block_indices):
        self.block_indices = np.asarray(block_indices)

    def __call__(self, vec, mat):
        from scipy.linalg import cho_factor, cho_solve

        new_vec = np.zeros_like(vec)
        for idx in self.block_indices:
            A = mat[np.ix_(idx, idx)]
            c = cho_factor(A, lower=True)
            new_vec[:, idx] = cho_solve(c, vec[:, idx].T).T
        return new_vec, np.sum(new_vec * vec, axis=-1, keepdims=True)


class DensePreconditioner:
    """``z = vec @ Pinv`` for a symmetric ``Pinv ~ A^-1`` (a Nystrom- / Cholesky-style preconditioner in the reference's
    protocol ``__call__(vec, mat) -> (z, rz)``, cggp/conjugate_gradient.py:125-128; the reference ships none for the
    matrix-free operator, this is the restatement of the new path's ``DensePreconditioner``)."""

    def __init__(self, pinv):
        self.pinv = np.asarray(pinv)

    def __call__(self, vec, mat):
        z = vec @ self.pinv
        return z, np.sum(z * vec, axis=-1, keepdims=True)


def conjugate_gradient(
    matrix,
    rhs,
    initial_solution,
    error_threshold,
    preconditioner: Optional[Callable] = None,
    max_iterations: Optional[int] = None,
    max_steps_cycle: int = 100,
    history: Optional[list] = None,
):
    """conjugate_gradient.py:24-122 (forward pass).  Returns ``(solution, (steps, 0.5*rz))``.

    ``history`` (extension, for the per-iteration parity check): if a list, ``0.5*|r_b|^2`` per RHS is appended at
    every evaluation of the stopping condition (so it has ``steps+1`` entries of shape [m]).
    """
    rhs = np.asarray(rhs)
    dtype = rhs.dtype
    matmul = matrix if callable(matrix) else (lambda V: V @ matrix)
    n = rhs.shape[-1]
    if preconditioner is None:  # :44-45
        preconditioner = EyePreconditioner()
    if max_iterations is None:  # :47-48
        max_iterations = n
    min_float = dtype.type(1e-16)  # :50 (1e-16 in the solution dtype, also for float32)
    zero = dtype.type(0.0)
    half = dtype.type(0.5)
    A = None if callable(matrix) else matrix
    b = rhs
    v = np.array(initial_solution, dtype=dtype, copy=True)

    def over_threshold(r):  # :59-62
        norm_r_sq = np.sum(np.square(r), axis=-1, keepdims=True)
        if history is not None:
            history.append((half * norm_r_sq)[:, 0].copy())
        return bool(np.any(half * norm_r_sq > error_threshold))

    # :87-92
    r = b - matmul(v)
    z, rz = preconditioner(r, A)
    p = z
    i = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        while over_threshold(r) and i < max_iterations:  # :93-95
            pA = matmul(p)  # :65
            denom = np.sum(p * pA, axis=-1, keepdims=True)  # :66
            gamma = rz / denom  # :67
            gamma = np.where(denom <= min_float, zero, gamma)  # :68
            v = v + gamma * p  # :69
            reset = i % max_steps_cycle == max_steps_cycle - 1  # :71 (pre-increment i)
            i += 1  # :70
            if reset:  # :72-76
                r = b - matmul(v)
            else:
                r = r - gamma * pA
            z, new_rz = preconditioner(r, A)  # :77
            z_update = p * new_rz / rz  # :78
            z_update = np.where(rz <= min_float, zero, z_update)  # :79
            p = z if reset else z + z_update  # :80-84
            rz = new_rz
    return v, (np.int32(i), half * rz)  # :96-98,120


def grad_conjugate_gradient(matrix, solution, dx, error_threshold, preconditioner=None,
                            max_iterations=None, max_steps_cycle=100):
    """conjugate_gradient.py:100-118: ``db = A^-1 dx`` by the same loop, ``dA = -solution^T @ db``."""
    db, _ = conjugate_gradient(matrix, dx, np.zeros_like(dx), error_threshold, preconditioner,
                               max_iterations, max_steps_cycle)
    dA = -solution.T @ db
    return dA, db


class ConjugateGradient:
    """conjugate_gradient.py:160-212: column-RHS adapter ``rhs [n, m] -> solution [n, m]``; drops stats."""

    def __init__(self, error_threshold, preconditioner=None, max_iterations=None, max_steps_cycle=None):
        self.error_threshold = error_threshold
        self.preconditioner = EyePreconditioner() if preconditioner is None else preconditioner
        self.max_iterations = max_iterations
        self.max_steps_cycle = max_steps_cycle
        self.last_stats = None  # extension: (steps, 0.5 rz) of the last call
        self.last_history = None

    def __call__(self, matrix, rhs, initial_solution=None):
        rhs = np.asarray(rhs).T  # :183
        initial_solution = np.zeros_like(rhs) if initial_solution is None else np.asarray(initial_solution).T
        max_iterations = self.max_iterations
        if max_iterations is None:
            max_iterations = rhs.shape[-1]  # :190-192 (= n)
        max_steps_cycle = self.max_steps_cycle
        if max_steps_cycle is None:
            max_steps_cycle = max_iterations + 1  # :194-196: never refresh
        hist = []
        solution, stats = conjugate_gradient(
            matrix, rhs, initial_solution, self.error_threshold, preconditioner=self.preconditioner,
            max_iterations=max_iterations, max_steps_cycle=max_steps_cycle, history=hist)
        self.last_stats = stats
        self.last_history = np.array(hist)
        return solution.T  # :211-212
