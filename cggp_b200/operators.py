"""Operator descriptors accepted at the ``matrix`` position of ``conjugate_gradient``:
a dense ``[n, n]`` tensor (the reference's only form, ``Kuu + Lambda``, cggp/models.py:301,337) or the matrix-free
north-star operator ``Sigma = Kuu + jitter I + s^-2 Kuf Kfu`` (the system GPflow's SGPR factorises, cggp/cli_utils.py:444-446)
whose data term is computed on the fly over this rank's shard of X and summed over ranks by one all-reduce per
iteration."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .kernels import (PreparedPoints, Stationary, kernel_matrix, kernel_matrix_param_grads, kuf_gram, prepare_points,
                      prepare_tf32, symm_matmul)


class LinearOperator:
    n: int
    dtype: torch.dtype
    device: torch.device

    def c_struct(self) -> _lib.Operator:  # pragma: no cover - interface
        raise NotImplementedError

    def matmul(self, V: torch.Tensor) -> torch.Tensor:  # pragma: no cover - interface
        raise NotImplementedError


class DenseOperator(LinearOperator):
    """``V @ A`` for a symmetric dense ``A`` (CG requires symmetry; the product reads rows of A, not columns)."""

    def __init__(self, A):
        A = _lib.row_major(_lib.as_device_tensor(A))
        if A.shape[0] != A.shape[1]:
            raise ValueError(f"matrix must be square, got {tuple(A.shape)}")
        self.A = A
        self.n = A.shape[0]
        self.dtype = A.dtype
        self.device = A.device

    def c_struct(self):
        op = _lib.Operator()
        op.type = _lib.OP_DENSE
        op.dtype = _lib.dtype_code(self.dtype)
        op.n = self.n
        op.dev_A = self.A.data_ptr()
        op.lda = self.A.stride(0)
        return op

    def matmul(self, V):
        V = _lib.row_major(V)
        ctx = _lib.context(self.device)
        ctx.use_current_stream()
        Y = torch.empty_like(V)
        ctx.check(ctx.lib.cggp_symm_matmul(ctx.handle, _lib.dtype_code(self.dtype), _lib.ptr(self.A),
                                           self.A.stride(0), self.n, _lib.ptr(V), V.stride(0), V.shape[0],
                                           _lib.ptr(Y), Y.stride(0)))
        return Y


class SGPROperator(LinearOperator):
    """``Sigma = Kuu + jitter I + Kuf Kfu / noise_variance`` applied matrix-free.

    ``X`` is THIS RANK's shard of the training inputs (rows); ``Z`` and ``Kuu`` are replicated.  With a communicator
    initialised on the context (``_lib.context().init_comm()``) every application all-reduces the partial product."""

    def __init__(self, kernel: Stationary, X, Z, noise_variance: float, jitter: float = 1e-6, variant: int = 0,
                 tf32_nsplit: int = 16):
        """``variant``: 0 auto, 1 two-sweep kernels, 3 float64 fused pipelined kernels, 4 float32 tensor-core (tcgen05 TF32)
        kernels - the default for float32 on sm_100 where the tiles fit (D <= 128).
        ``tf32_nsplit`` selects the tensor-core arithmetic: 16 (default) = 3xFP16 with a per-row power-of-two scale
        (float32-accurate distances at twice the TF32 rate), 3 = 3xTF32 (same accuracy), 1 = one TF32 pass (~1e-3
        relative on the distances)."""
        self.kernel = kernel
        self._X, self._Z = X, Z
        self._snapshot()
        self.PZ = Z if isinstance(Z, PreparedPoints) else prepare_points(Z, self._lengthscales)
        self.PX = X if isinstance(X, PreparedPoints) else prepare_points(X, self._lengthscales, self.PZ.P.dtype)
        # a tensor that requires grad makes the solve differentiable in the likelihood variance (_SGPRSolveFn)
        self._noise_t = noise_variance if isinstance(noise_variance, torch.Tensor) and noise_variance.requires_grad \
            else None
        self.noise_variance = float(noise_variance.detach()) if isinstance(noise_variance, torch.Tensor) \
            else float(noise_variance)
        self.jitter = float(jitter)
        self.variant = int(variant)
        self.n = self.PZ.n
        self.dtype = self.PZ.P.dtype
        self.device = self.PZ.P.device
        self.tf32_nsplit = int(tf32_nsplit)
        self._prepare_derived()

    # An operator is valid for ONE value of the kernel hyper-parameters: the scaled points, Kuu and the tensor-core
    # arrays are prepared from a snapshot taken at construction (or at the last `refresh()`), and every product /
    # gradient uses that snapshot.  After an in-place optimiser step on the kernel's parameters the operator is
    # `stale`; using it raises instead of silently mixing old points with new variances.
    def _snapshot(self):
        self._variance = float(self.kernel.variance)
        self._lengthscales = self.kernel.lengthscales.clone()

    @property
    def stale(self) -> bool:
        return self._variance != float(self.kernel.variance) or \
            not torch.equal(self._lengthscales, self.kernel.lengthscales)

    def _check_fresh(self):
        if self.stale:
            raise _lib.CggpError("SGPROperator is stale: the kernel hyper-parameters changed since it was prepared; "
                                 "call operator.refresh() (an operator is valid for one parameter value)")

    def refresh(self, noise_variance=None):
        """Re-prepare the scaled points, Kuu and the tensor-core arrays from the kernel's CURRENT hyper-parameters
        (and optionally a new likelihood variance)."""
        self._snapshot()
        if noise_variance is not None:
            self._noise_t = noise_variance if isinstance(noise_variance, torch.Tensor) and \
                noise_variance.requires_grad else None
            self.noise_variance = float(noise_variance.detach()) if isinstance(noise_variance, torch.Tensor) \
                else float(noise_variance)
        if isinstance(self._Z, PreparedPoints) or isinstance(self._X, PreparedPoints):
            raise _lib.CggpError("an operator built from prepared points cannot be refreshed: rebuild it")
        self.PZ = prepare_points(self._Z, self._lengthscales)
        self.PX = prepare_points(self._X, self._lengthscales, self.PZ.P.dtype)
        self._prepare_derived()
        return self

    def _prepare_derived(self):
        self.Kuu = kernel_matrix(self.kernel.kind, self._variance, self.PZ, self.PZ, jitter=self.jitter)
        self.X32 = self.Z32 = None
        if self.dtype == torch.float32 and self.variant in (0, 4):
            ctx = _lib.context(self.device)
            if ctx.lib.cggp_tf32_supported(ctx.handle, self.PZ.D, self.tf32_nsplit):
                self.X32 = prepare_tf32(self.PX, self.tf32_nsplit)
                self.Z32 = prepare_tf32(self.PZ, self.tf32_nsplit)
            elif self.variant == 4:
                raise _lib.CggpError("the tcgen05 path needs sm_100 and D <= 128 (3xFP16, 3xTF32) / 256 (1xTF32)")

    def c_struct(self):
        self._check_fresh()
        op = _lib.Operator()
        op.type = _lib.OP_SGPR
        op.dtype = _lib.dtype_code(self.dtype)
        op.n = self.n
        op.dev_A = self.Kuu.data_ptr()
        op.lda = self.Kuu.stride(0)
        op.kind = self.kernel.kind
        op.D = self.PZ.D
        op.variance = self._variance
        op.scale = 1.0 / self.noise_variance
        op.dev_PX = self.PX.P.data_ptr()
        op.dev_normsX = self.PX.norms.data_ptr()
        op.n_local = self.PX.n
        op.dev_PZ = self.PZ.P.data_ptr()
        op.dev_normsZ = self.PZ.norms.data_ptr()
        op.ldp = self.PZ.ldp
        op.variant = self.variant if self.variant != 4 else 0
        if self.X32 is not None:
            op.dev_X32_big, op.dev_X32_small = self.X32.big.data_ptr(), self.X32.small.data_ptr()
            op.dev_x32_norms = self.X32.norms.data_ptr()
            op.dev_Z32_big, op.dev_Z32_small = self.Z32.big.data_ptr(), self.Z32.small.data_ptr()
            op.dev_z32_norms = self.Z32.norms.data_ptr()
            op.tf32_nsplit = self.tf32_nsplit
        return op

    def kuf_kfu_matmul(self, V, variant=None, allreduce=True):
        """``V @ (Kuf Kfu)`` over the local shard (+ all-reduce)."""
        self._check_fresh()
        V = _lib.row_major(V)
        ctx = _lib.context(self.device)
        ctx.use_current_stream()
        W = torch.empty_like(V)
        use = self.variant if variant is None else int(variant)
        if self.X32 is not None and use in (0, 4):
            ctx.check(ctx.lib.cggp_kuf_kfu_matvec_tf32(
                ctx.handle, self.kernel.kind, self._variance, _lib.ptr(self.X32.big), _lib.ptr(self.X32.small),
                _lib.ptr(self.X32.norms), self.PX.n, _lib.ptr(self.Z32.big), _lib.ptr(self.Z32.small),
                _lib.ptr(self.Z32.norms), self.n, self.PZ.D, _lib.ptr(V), V.stride(0), V.shape[0], _lib.ptr(W),
                W.stride(0), self.tf32_nsplit))
            if allreduce and ctx.world > 1:
                ctx.allreduce_sum_(W)
            return W
        if use == 4:
            raise _lib.CggpError("variant 4 (tcgen05 TF32) is not available for this operator")
        ctx.check(ctx.lib.cggp_kuf_kfu_matvec(
            ctx.handle, _lib.dtype_code(self.dtype), self.kernel.kind, self._variance,
            _lib.ptr(self.PX.P), _lib.ptr(self.PX.norms), self.PX.n, _lib.ptr(self.PZ.P), _lib.ptr(self.PZ.norms),
            self.n, self.PZ.D, self.PZ.ldp, _lib.ptr(V), V.stride(0), V.shape[0], _lib.ptr(W), W.stride(0),
            use))
        if allreduce and ctx.world > 1:
            ctx.allreduce_sum_(W)
        return W

    def matmul(self, V):
        W = self.kuf_kfu_matmul(V)
        return DenseOperator(self.Kuu).matmul(V) + W / self.noise_variance

    @property
    def trainable(self) -> bool:
        return self.kernel.trainable or self._noise_t is not None

    def solve_param_grads(self, sol, lam):
        """Hyper-parameter gradients of a loss through the solve ``sol Sigma = rhs``, given ``lam = dL/dsol Sigma^-1``
        (the adjoint solve): ``dL/dSigma = -sol^T lam`` (cggp/conjugate_gradient.py:117), pulled back through
        ``Sigma = Kuu + Kuf Kfu / s2`` without forming anything of size N x M at once:
        ``dL/dKuf = -(sol^T (lam Kuf) + lam^T (sol Kuf)) / s2`` chunk by chunk into ``cggp_kernel_matrix_backward``,
        ``dL/ds2 = sum_b <sol_b Kuf, lam_b Kuf> / s2^2``.  The shard terms are all-reduced.  Returns
        ``(dL/dvariance [1], dL/dlengthscales [D], dL/ds2 [1])``."""
        self._check_fresh()
        kind, var, ls = self.kernel.kind, self._variance, self._lengthscales
        D, M = self.PZ.D, self.n
        g = kernel_matrix_param_grads(kind, var, ls, self.PZ, self.PZ, -(sol.t() @ lam))  # replicated Kuu term
        shard = torch.zeros((2 + D,), dtype=self.dtype, device=self.device)
        inv = 1.0 / self.noise_variance
        step = max(1, (1 << 27) // max(M, 1))
        for s in range(0, self.PX.n, step):
            rows = self.PX.rows(s, min(self.PX.n, s + step))
            Kc = kernel_matrix(kind, var, self.PZ, rows)
            Ts, Tl = sol @ Kc, lam @ Kc
            dK = torch.addmm(sol.t() @ Tl, lam.t(), Ts).mul_(-inv)
            shard[: 1 + D] += kernel_matrix_param_grads(kind, var, ls, self.PZ, rows, dK)
            shard[1 + D] += (Ts * Tl).sum()
        ctx = _lib.context(self.device)
        if ctx.world > 1:
            ctx.allreduce_sum_(shard)
        g = g + shard[: 1 + D]
        return g[:1], g[1:], shard[1 + D:] * (inv * inv)

    def kuf_times(self, Y):
        """``Kuf @ Y`` for the local shard ``Y [n_local, P]`` (+ all-reduce): the right-hand side ``Kuf y``.
        Computed in row batches so that nothing of size N x M is ever resident."""
        self._check_fresh()
        Y = _lib.row_major(_lib.as_device_tensor(Y, self.dtype))
        ctx = _lib.context(self.device)
        if self.dtype == torch.float64 and self.PZ.D <= 31 and self.n <= 256 * 148 and self.variant in (0, 3):
            # fused: phase 2 of the pipelined kernel with the row weights given (cggp_kuf_times)
            ctx.use_current_stream()
            W = torch.empty((Y.shape[1], self.n), dtype=self.dtype, device=self.device)
            ctx.check(ctx.lib.cggp_kuf_times(
                ctx.handle, _lib.dtype_code(self.dtype), self.kernel.kind, self._variance, _lib.ptr(self.PX.P),
                _lib.ptr(self.PX.norms), self.PX.n, _lib.ptr(self.PZ.P), _lib.ptr(self.PZ.norms), self.n, self.PZ.D,
                self.PZ.ldp, _lib.ptr(Y), Y.stride(0), Y.shape[1], _lib.ptr(W), W.stride(0)))
            out = W.t().contiguous()
            if ctx.world > 1:
                ctx.allreduce_sum_(out)
            return out
        if self.X32 is not None and self.variant in (0, 4):
            # float32: one gram-contraction sweep on the tensor cores (rows = Z, columns = X, weights = Y^T)
            ctx.use_current_stream()
            Yt = Y.t().contiguous()
            W = torch.empty((Y.shape[1], self.n), dtype=self.dtype, device=self.device)
            ctx.check(ctx.lib.cggp_kuf_times_tf32(
                ctx.handle, self.kernel.kind, self._variance, _lib.ptr(self.X32.big), _lib.ptr(self.X32.small),
                _lib.ptr(self.X32.norms), self.PX.n, _lib.ptr(self.Z32.big), _lib.ptr(self.Z32.small),
                _lib.ptr(self.Z32.norms), self.n, self.PZ.D, _lib.ptr(Yt), Yt.stride(0), Yt.shape[0], _lib.ptr(W),
                W.stride(0), self.tf32_nsplit))
            out = W.t().contiguous()
            if ctx.world > 1:
                ctx.allreduce_sum_(out)
            return out
        out = torch.zeros((self.n, Y.shape[1]), dtype=self.dtype, device=self.device)
        step = max(1, (1 << 27) // max(self.n, 1))
        for s in range(0, self.PX.n, step):
            e = min(self.PX.n, s + step)
            Kzx = kernel_matrix(self.kernel.kind, self._variance, self.PZ, self.PX.rows(s, e))
            out += Kzx @ Y[s:e]
        ctx = _lib.context(self.device)
        if ctx.world > 1:
            ctx.allreduce_sum_(out)
        return out


    def gram(self, rows: PreparedPoints = None) -> torch.Tensor:
        """``Kuf Kfu`` as a dense ``[M, M]`` matrix over the local shard (+ all-reduce when ``rows`` is None): the
        materialised form GPflow's SGPR builds (``A A^T``), needed where a log-determinant or a trace is required
        (``SGPR.elbo``).  Row chunks of ``Kuf`` are evaluated on the fly by ``cggp_kernel_matrix``; the rank-k updates
        run on the library's own FP64 DMMA GEMM, lower tiles only (``cggp_kuf_gram``)."""
        self._check_fresh()
        src = self.PX if rows is None else rows
        G = kuf_gram(self.kernel.kind, self._variance, self.PZ, src)
        ctx = _lib.context(self.device)
        if rows is None and ctx.world > 1:
            ctx.allreduce_sum_(G)
        return G

    def nystrom_preconditioner(self, num_rows: int = None, seed: int = 0):
        """Preconditioner for ``Sigma``: ``P = Kuu + jitter I + (N / n_s) Kuf_s Kfu_s / noise`` from a uniform
        subsample of ``n_s`` training rows (default ``4 M`` over all ranks), inverted once through a Cholesky
        factorisation (setup cost ``n_s M^2 + M^3``; library factorisation, outside the hot loop).  Every rank draws
        its share of the subsample from its own shard; the ``[M, M]`` Gram matrix and the row counts are all-reduced,
        so all ranks hold the same ``Pinv``.  Returns a ``DensePreconditioner``."""
        from .conjugate_gradient import DensePreconditioner

        ctx = _lib.context(self.device)
        world = ctx.world
        n_loc = self.PX.n
        want = (4 * self.n if num_rows is None else int(num_rows))
        take = min(n_loc, max(1, want // world))
        gen = torch.Generator(device="cpu").manual_seed(seed + 7919 * ctx.rank)
        idx = torch.randperm(n_loc, generator=gen)[:take].sort().values.to(self.device)
        sub = PreparedPoints(self.PX.P[idx].contiguous(), self.PX.norms[idx].contiguous(), self.PX.D)
        gram = self.gram(sub)
        counts = torch.tensor([float(take), float(n_loc)], dtype=self.dtype, device=self.device)
        if world > 1:
            ctx.allreduce_sum_(gram)
            ctx.allreduce_sum_(counts)
        scale = float(counts[1] / counts[0]) / self.noise_variance
        P = self.Kuu + scale * gram
        L = torch.linalg.cholesky(P)
        return DensePreconditioner(torch.cholesky_inverse(L))


class _SGPRSolveFn(torch.autograd.Function):
    """Differentiable matrix-free solve ``sol Sigma = rhs`` on an ``SGPROperator``: forward = the device CG loop,
    backward = a second CG solve with the incoming gradient as right-hand side (the reference's custom gradient,
    cggp/conjugate_gradient.py:100-118), ``dL/drhs = lam``, and ``dL/dSigma = -sol^T lam`` pulled back to the kernel
    hyper-parameters and the likelihood variance by ``SGPROperator.solve_param_grads``."""

    @staticmethod
    def forward(ctx, variance_t, lengthscales_t, noise_t, rhs, op, v0, cfg):
        from .conjugate_gradient import _solve

        sol, steps, err, hist = _solve(op, rhs.detach(), None if v0 is None else v0.detach(), *cfg)
        ctx.save_for_backward(sol)
        ctx.op, ctx.cfg = op, cfg
        ctx.shapes = (variance_t, lengthscales_t, noise_t)
        ctx.stats = (steps, err, hist)
        ctx.mark_non_differentiable(err)
        return sol, err

    @staticmethod
    def backward(ctx, dx, _derr):
        from .conjugate_gradient import _solve

        (sol,) = ctx.saved_tensors
        op = ctx.op
        lam, _, _, _ = _solve(op, dx.contiguous(), None, *ctx.cfg)
        variance_t, lengthscales_t, noise_t = ctx.shapes
        g_var = g_ls = g_noise = None
        if op.trainable:
            gv, gl, gn = op.solve_param_grads(sol, lam)
            if lengthscales_t.numel() == 1:
                gl = gl.sum()
            g_var = gv.reshape(variance_t.shape).to(variance_t.dtype)
            g_ls = gl.reshape(lengthscales_t.shape).to(lengthscales_t.dtype)
            g_noise = gn.reshape(noise_t.shape).to(noise_t.dtype)
        return g_var, g_ls, g_noise, lam, None, None, None


def as_operator(matrix) -> LinearOperator:
    return matrix if isinstance(matrix, LinearOperator) else DenseOperator(matrix)
