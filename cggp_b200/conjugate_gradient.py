"""Drop-in mirror of ``cggp/conjugate_gradient.py`` on the B200 path.

Same names, positional order and return values as the reference:

* ``conjugate_gradient(matrix, rhs, initial_solution, error_threshold, preconditioner=None, max_iterations=None,
  max_steps_cycle=100) -> (solution [m, n], (steps int32, 0.5*rz [m, 1]))``   (conjugate_gradient.py:24-32,120)
* ``ConjugateGradient(error_threshold, preconditioner=None, max_iterations=None, max_steps_cycle=None)(matrix,
  rhs [n, m], initial_solution=None) -> solution [n, m]``                      (conjugate_gradient.py:160-212)
* ``EyePreconditioner`` / ``BlockPreconditioner(block_indices)``               (conjugate_gradient.py:125-157)

Extensions, at the same argument positions: ``matrix`` may be a ``LinearOperator`` (``operators.py``); tensors may
be anything exposing ``__dlpack__``.  The whole loop runs on the device inside ``cggp_cg_solve`` (C ABI); like the
reference's ``tf.custom_gradient`` the solve is differentiable (backward = a second CG solve, ``dA = -solution^T db``,
conjugate_gradient.py:100-118): for dense matrices as in the reference, and for the matrix-free ``SGPROperator`` with
``dA`` pulled back to the kernel hyper-parameters / likelihood variance chunk by chunk (``_SGPRSolveFn``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Union

import torch

from . import _lib
from .operators import DenseOperator, LinearOperator, SGPROperator, _SGPRSolveFn, as_operator

Tensor = torch.Tensor


class CGPreconditioner:
    def __call__(self, vec: Tensor, mat) -> Tuple[Tensor, Tensor]:
        raise NotImplementedError

    def c_struct(self, operator: LinearOperator):
        raise NotImplementedError


class EyePreconditioner(CGPreconditioner):
    """``z = r``, ``rz = sum r^2`` (conjugate_gradient.py:131-134); fused into the CG step kernel."""

    def __call__(self, vec, mat=None):
        return vec, (vec * vec).sum(-1, keepdim=True)

    def c_struct(self, operator):
        return None, ()


class BlockPreconditioner(CGPreconditioner):
    """Block-Jacobi ``z[:, blk] = A[blk, blk]^-1 r[:, blk]`` (the evident intent of conjugate_gradient.py:137-157;
    as written the reference gathers RHS rows instead of vector elements, line 144, and is never instantiated).
    ``block_indices``: int tensor ``[num_blocks, block_size]`` partitioning ``range(n)``; blocks are factorised on
    the device (``cggp_block_cholesky``) once per solve and applied inside the fused step kernel."""

    def __init__(self, block_indices) -> None:
        self.block_indices = torch.as_tensor(block_indices, dtype=torch.int64)
        if self.block_indices.dim() != 2:
            raise ValueError("block_indices must be [num_blocks, block_size]")

    def _validate(self, n):
        flat = self.block_indices.reshape(-1).cpu()
        if flat.numel() != n or not torch.equal(torch.sort(flat).values, torch.arange(n)):
            raise ValueError("block_indices must partition range(n)")

    def factorise(self, operator: LinearOperator):
        if isinstance(operator, DenseOperator):
            A = operator.A
        elif hasattr(operator, "dense_block_source"):
            A = operator.dense_block_source()
        else:
            raise TypeError("BlockPreconditioner needs dense access to the diagonal blocks")
        n = operator.n
        self._validate(n)
        idx = self.block_indices.to(A.device).contiguous()
        nb, bs = idx.shape
        chol = torch.empty((nb, bs, bs), dtype=A.dtype, device=A.device)
        ctx = _lib.context(A.device)
        ctx.use_current_stream()
        ctx.check(ctx.lib.cggp_block_cholesky(ctx.handle, _lib.dtype_code(A.dtype), _lib.ptr(A), A.stride(0), n,
                                              _lib.ptr(idx), nb, bs, _lib.ptr(chol)))
        return idx, chol

    def c_struct(self, operator):
        idx, chol = self.factorise(operator)
        pc = _lib.Precond()
        pc.type = _lib.PRECOND_BLOCK
        pc.num_blocks, pc.block_size = idx.shape
        pc.dev_block_indices = idx.data_ptr()
        pc.dev_chol = chol.data_ptr()
        return pc, (idx, chol)

    def __call__(self, vec, mat):
        pc, keep = self.c_struct(as_operator(mat))
        vec = _lib.row_major(_lib.as_device_tensor(vec)).contiguous()
        z = torch.empty_like(vec)
        ctx = _lib.context(vec.device)
        ctx.use_current_stream()
        ctx.check(ctx.lib.cggp_block_precond_apply(ctx.handle, _lib.dtype_code(vec.dtype), vec.shape[0], vec.shape[1],
                                                   _lib.ptr(vec), C.byref(pc), _lib.ptr(z)))
        del keep
        return z, (z * vec).sum(-1, keepdim=True)


class DensePreconditioner(CGPreconditioner):
    """``z = r @ Pinv`` with a symmetric ``Pinv ~ A^-1`` resident on the device (``CGGP_PRECOND_DENSE``): the
    Nystrom- / Cholesky-style preconditioner of the matrix-free operator (``SGPROperator.nystrom_preconditioner``).
    Follows the reference's protocol ``__call__(vec [m, n], mat) -> (z [m, n], rz [m, 1])``
    (conjugate_gradient.py:125-128); inside the device loop the product is one symmetric GEMV/GEMM per iteration."""

    def __init__(self, pinv) -> None:
        pinv = _lib.row_major(_lib.as_device_tensor(pinv))
        if pinv.dim() != 2 or pinv.shape[0] != pinv.shape[1]:
            raise ValueError("pinv must be a square matrix")
        self.pinv = pinv

    def c_struct(self, operator):
        if self.pinv.shape[0] != operator.n or self.pinv.dtype != operator.dtype:
            raise ValueError("preconditioner does not match the operator (size / dtype)")
        pc = _lib.Precond()
        pc.type = _lib.PRECOND_DENSE
        pc.dev_pinv = self.pinv.data_ptr()
        pc.ldpinv = self.pinv.stride(0)
        return pc, (self.pinv,)

    def __call__(self, vec, mat=None):
        z = DenseOperator(self.pinv).matmul(_lib.row_major(vec))
        return z, (z * vec).sum(-1, keepdim=True)


def _solve(operator: LinearOperator, rhs: Tensor, initial_solution: Optional[Tensor], error_threshold: float,
           preconditioner: Optional[CGPreconditioner], max_iterations: int, max_steps_cycle: int,
           want_history: bool, check_every: int = 16):
    rhs = _lib.row_major(rhs).contiguous()
    B, n = rhs.shape
    if n != operator.n:
        raise ValueError(f"rhs has {n} columns but the system size is {operator.n}")
    if rhs.dtype != operator.dtype:
        raise TypeError(f"rhs dtype {rhs.dtype} != operator dtype {operator.dtype}")
    ctx = _lib.context(rhs.device)
    ctx.use_current_stream()
    x0 = None
    if initial_solution is not None:
        x0 = _lib.row_major(initial_solution).contiguous()
        if x0.shape != rhs.shape:
            raise ValueError("initial_solution must have the shape of rhs")
    solution = torch.empty_like(rhs)
    half_rz = torch.empty((B,), dtype=rhs.dtype, device=rhs.device)
    history = None
    cap = 0
    if want_history:
        cap = int(max_iterations) + 1
        history = torch.full((cap, B), float("nan"), dtype=rhs.dtype, device=rhs.device)
    if preconditioner is None:
        preconditioner = EyePreconditioner()
    pc, keep = preconditioner.c_struct(operator)
    op = operator.c_struct()
    steps = C.c_int32(0)
    ctx.check(ctx.lib.cggp_cg_solve(
        ctx.handle, C.byref(op), _lib.ptr(rhs), _lib.ptr(x0), B, float(error_threshold), int(max_iterations),
        int(max_steps_cycle), C.byref(pc) if pc is not None else None, int(check_every), _lib.ptr(solution),
        C.byref(steps), _lib.ptr(half_rz), _lib.ptr(history), cap))
    del keep
    if history is not None:
        history = history[: steps.value + 1]
    return solution, steps.value, half_rz.reshape(B, 1), history


class _CGFunction(torch.autograd.Function):
    """Forward = device CG; backward = conjugate_gradient.py:100-118 (second CG solve with rhs = dx)."""

    @staticmethod
    def forward(ctx, A, b, v0, cfg):
        sol, steps, err, hist = _solve(DenseOperator(A.detach()), b.detach(), None if v0 is None else v0.detach(), *cfg)
        ctx.save_for_backward(A.detach(), sol)
        ctx.cfg = cfg
        ctx.stats = (steps, err, hist)
        ctx.mark_non_differentiable(err)
        return sol, err

    @staticmethod
    def backward(ctx, dx, _derr):
        A, sol = ctx.saved_tensors
        db, _, _, _ = _solve(DenseOperator(A), dx.contiguous(), None, *ctx.cfg)
        dA = -(sol.t() @ db)  # :117
        return dA, db, None, None  # :118 (no gradient to the initial solution)


def conjugate_gradient(
    matrix,
    rhs,
    initial_solution,
    error_threshold: float,
    preconditioner: Optional[CGPreconditioner] = None,
    max_iterations: Optional[int] = None,
    max_steps_cycle: int = 100,
    *,
    return_history: bool = False,
):
    """Conjugate gradient for ``v A = b`` with ``rhs`` rows as independent right-hand sides
    (conjugate_gradient.py:24-122).  Returns ``(solution, (steps, 0.5*rz))``; with ``return_history=True`` a third
    stats entry holds ``0.5*|r_b|^2`` at every evaluation of the stopping condition, shape ``[steps+1, m]``."""
    rhs_t = _lib.as_device_tensor(rhs)
    x0_t = None if initial_solution is None else _lib.as_device_tensor(initial_solution, rhs_t.dtype)
    dense = not isinstance(matrix, LinearOperator)
    if dense:
        matrix = _lib.as_device_tensor(matrix, rhs_t.dtype)
        n = matrix.shape[0]
    else:
        n = matrix.n
    if max_iterations is None:
        max_iterations = n  # :47-48
    cfg = (float(error_threshold), preconditioner, int(max_iterations), int(max_steps_cycle), bool(return_history))
    needs_grad = dense and torch.is_grad_enabled() and (matrix.requires_grad or rhs_t.requires_grad)
    op_grad = (not dense) and isinstance(matrix, SGPROperator) and torch.is_grad_enabled() and \
        (matrix.trainable or rhs_t.requires_grad)
    if needs_grad:
        sol, err = _CGFunction.apply(matrix, rhs_t, x0_t, cfg)
        # stats of the forward pass are attached by the Function
        steps, _, hist = sol.grad_fn.stats if sol.grad_fn is not None and hasattr(sol.grad_fn, "stats") else (-1, None, None)
    elif op_grad:
        # matrix-free operator with trainable hyper-parameters (or a right-hand side that needs a gradient)
        var_t, ls_t = matrix.kernel.param_tensors(rhs_t.dtype, rhs_t.device)
        noise_t = matrix._noise_t if matrix._noise_t is not None else \
            torch.tensor(matrix.noise_variance, dtype=rhs_t.dtype, device=rhs_t.device)
        sol, err = _SGPRSolveFn.apply(var_t, ls_t, noise_t, rhs_t, matrix, x0_t, cfg)
        steps, _, hist = sol.grad_fn.stats if sol.grad_fn is not None and hasattr(sol.grad_fn, "stats") else (-1, None, None)
    else:
        sol, steps, err, hist = _solve(as_operator(matrix), rhs_t, x0_t, *cfg)
    stats_steps = torch.tensor(steps, dtype=torch.int32)
    if return_history:
        return sol, (stats_steps, err, hist)
    return sol, (stats_steps, err)


class ConjugateGradient:
    """Config holder + column-RHS adapter (conjugate_gradient.py:160-212): ``rhs [n, m] -> solution [n, m]``;
    ``max_iterations=None -> n``; ``max_steps_cycle=None -> max_iterations + 1`` (never refresh); stats are dropped by
    the reference and kept here in ``last_stats`` / ``last_history`` (set ``record_history=True``)."""

    def __init__(
        self,
        error_threshold: Union[Tensor, float],
        preconditioner: Optional[CGPreconditioner] = None,
        max_iterations: Optional[int] = None,
        max_steps_cycle: Optional[int] = None,
        *,
        record_history: bool = False,
    ):
        self.error_threshold = error_threshold
        if preconditioner is None:
            preconditioner = EyePreconditioner()
        self.preconditioner = preconditioner
        self.max_iterations = max_iterations
        self.max_steps_cycle = max_steps_cycle
        self.record_history = record_history
        self.last_stats = None
        self.last_history = None

    def __call__(self, matrix, rhs, initial_solution=None) -> Tensor:
        rhs = _lib.as_device_tensor(rhs).t()  # :183
        if initial_solution is not None:
            initial_solution = _lib.as_device_tensor(initial_solution).t()  # :188
        n = matrix.n if isinstance(matrix, LinearOperator) else matrix.shape[-1]
        max_iterations = self.max_iterations
        if max_iterations is None:
            max_iterations = n  # :190-192
        max_steps_cycle = self.max_steps_cycle
        if max_steps_cycle is None:
            max_steps_cycle = max_iterations + 1  # :194-196
        out = conjugate_gradient(
            matrix, rhs, initial_solution, float(self.error_threshold), preconditioner=self.preconditioner,
            max_iterations=max_iterations, max_steps_cycle=max_steps_cycle, return_history=self.record_history)
        solution, stats = out
        self.last_stats = stats[:2]
        self.last_history = stats[2] if self.record_history else None
        return solution.t()  # :211-212
