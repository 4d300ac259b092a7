"""Stationary kernels with the GPflow surface the reference uses (``kernel(X)``, ``kernel(X, X2)``,
``kernel(X, full_cov=False)``, ``.K``, ``.K_diag``, ``.variance``, ``.lengthscales``), evaluated on the GPU
through the C ABI (``cggp_prepare_points`` + ``cggp_kernel_matrix``).

Replaces the GPflow calls at ``cggp/models.py:300,333-335``, ``cggp/distance.py:17-20,26-29`` and the kernel
factory of ``cggp/cli_utils.py:363-368`` (default ``Matern32(variance=1, lengthscales=[1]*D)``).  Arithmetic follows
GPflow: inputs divided by the lengthscales, expanded squared distance without clamp, ``max(r2, 1e-36)`` before the
square root for the Matern family.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class PreparedPoints:
    """Scaled rows ``X / lengthscales`` zero-padded to ``ldp`` columns, and their squared norms."""

    P: torch.Tensor      # [n, ldp]
    norms: torch.Tensor  # [n]
    D: int

    @property
    def n(self) -> int:
        return self.P.shape[0]

    @property
    def ldp(self) -> int:
        return self.P.shape[1]

    def rows(self, start: int, stop: int) -> "PreparedPoints":
        return PreparedPoints(self.P[start:stop], self.norms[start:stop], self.D)


def prepare_points(X, lengthscales, dtype=None) -> PreparedPoints:
    X = _lib.row_major(_lib.as_device_tensor(X, dtype))
    ctx = _lib.context(X.device)
    ctx.use_current_stream()
    n, D = X.shape
    ldp = int(ctx.lib.cggp_prepared_ld(D))
    P = torch.empty((n, ldp), dtype=X.dtype, device=X.device)
    norms = torch.empty((n,), dtype=X.dtype, device=X.device)
    ls = torch.as_tensor(lengthscales, dtype=torch.float64).reshape(-1).cpu()
    if ls.numel() not in (1, D):
        raise ValueError(f"lengthscales must have 1 or {D} entries, got {ls.numel()}")
    arr = (C.c_double * ls.numel())(*ls.tolist())
    ctx.check(ctx.lib.cggp_prepare_points(ctx.handle, _lib.dtype_code(X.dtype), _lib.ptr(X), n, D, X.stride(0),
                                          arr, ls.numel(), _lib.ptr(P), ldp, _lib.ptr(norms)))
    return PreparedPoints(P, norms, D)


@dataclass
class TF32Points:
    """Prepared float32 points in the layout the tensor cores read (``cggp_tf32_prepare``): canonical K-major chunks,
    rows padded to 128; ``big`` = the interleaved arrays the products stream (TF32 big [| small], FP16 H | L),
    ``small`` = the row-role arrays of the 3xFP16 mode; norms padded with zeros."""

    big: torch.Tensor
    small: torch.Tensor
    norms: torch.Tensor
    n: int
    D: int
    nsplit: int = 3


def prepare_tf32(points: PreparedPoints, nsplit: int = 3) -> TF32Points:
    """Tensor-core operand arrays of prepared float32 points for ``nsplit`` = 3 (3xTF32), 1 (one TF32 pass) or 16
    (3xFP16)."""
    import ctypes as C

    if points.P.dtype != torch.float32:
        raise TypeError("the tensor-core path is for float32 points")
    ctx = _lib.context(points.P.device)
    ctx.use_current_stream()
    rows = int(ctx.lib.cggp_tf32_rows(points.n))
    ns, nr = C.c_int64(0), C.c_int64(0)
    ctx.check(ctx.lib.cggp_tf32_sizes(int(nsplit), points.n, points.D, C.byref(ns), C.byref(nr)))
    dev = points.P.device
    big = torch.empty((max(ns.value, 1),), dtype=torch.float32, device=dev)
    small = torch.empty((max(nr.value, 1),), dtype=torch.float32, device=dev)
    norms = torch.empty((max(rows, 1),), dtype=torch.float32, device=dev)
    P = points.P if points.P.stride(1) == 1 else points.P.contiguous()
    ctx.check(ctx.lib.cggp_tf32_prepare(ctx.handle, int(nsplit), _lib.ptr(P), _lib.ptr(points.norms.contiguous()),
                                        points.n, points.D, P.stride(0), _lib.ptr(big), _lib.ptr(small),
                                        _lib.ptr(norms)))
    return TF32Points(big, small, norms, points.n, points.D, int(nsplit))


def kernel_matrix(kind, variance, A: PreparedPoints, B: PreparedPoints, *, output=_lib.OUT_KERNEL,
                  distance=_lib.DIST_EUCLIDEAN, jitter=0.0, out=None) -> torch.Tensor:
    if A.D != B.D or A.ldp != B.ldp or A.P.dtype != B.P.dtype:
        raise ValueError("point sets disagree in dimension / dtype")
    ctx = _lib.context(A.P.device)
    ctx.use_current_stream()
    if out is None:
        out = torch.empty((A.n, B.n), dtype=A.P.dtype, device=A.P.device)
    step = 65535 * 64  # rows per call (grid.y limit)
    for s in range(0, max(A.n, 1), step):
        e = min(A.n, s + step)
        if e <= s:
            break
        ctx.check(ctx.lib.cggp_kernel_matrix(
            ctx.handle, _lib.dtype_code(A.P.dtype), int(kind), float(variance), int(output), int(distance),
            _lib.ptr(A.P[s:e]), _lib.ptr(A.norms[s:e]), e - s, _lib.ptr(B.P), _lib.ptr(B.norms), B.n, A.D, A.ldp,
            float(jitter), _lib.ptr(out[s:e]), out.stride(0)))
    return out


def kuf_gram(kind, variance, PZ: PreparedPoints, PX: PreparedPoints, out=None, accumulate=False) -> torch.Tensor:
    """``Kuf Kfu`` ``[M, M]`` over the rows of ``PX`` (``cggp_kuf_gram``: ``Kuf`` evaluated in L2-sized row chunks,
    contracted by the library's FP64 DMMA GEMM as a symmetric rank-k update).  Not all-reduced."""
    ctx = _lib.context(PZ.P.device)
    ctx.use_current_stream()
    M = PZ.n
    if out is None:
        out = torch.empty((M, M), dtype=PZ.P.dtype, device=PZ.P.device)
        accumulate = False
    ctx.check(ctx.lib.cggp_kuf_gram(
        ctx.handle, _lib.dtype_code(PZ.P.dtype), int(kind), float(variance), _lib.ptr(PX.P), _lib.ptr(PX.norms), PX.n,
        _lib.ptr(PZ.P), _lib.ptr(PZ.norms), M, PZ.D, PZ.ldp, _lib.ptr(out), out.stride(0), 1 if accumulate else 0))
    return out


def symm_matmul(A: torch.Tensor, V: torch.Tensor) -> torch.Tensor:
    """``V @ A`` for a symmetric ``A`` through ``cggp_symm_matmul`` (HBM-bound GEMV for few rows of V, FP64 DMMA GEMM
    for many)."""
    A, V = _lib.row_major(A), _lib.row_major(V)
    ctx = _lib.context(A.device)
    ctx.use_current_stream()
    Y = torch.empty_like(V)
    if V.shape[0] and A.shape[0]:
        ctx.check(ctx.lib.cggp_symm_matmul(ctx.handle, _lib.dtype_code(A.dtype), _lib.ptr(A), A.stride(0), A.shape[0],
                                           _lib.ptr(V), V.stride(0), V.shape[0], _lib.ptr(Y), Y.stride(0)))
    return Y


def kernel_matrix_param_grads(kind, variance, lengthscales, A: PreparedPoints, B: PreparedPoints, G) -> torch.Tensor:
    """``[dL/dvariance, dL/dlengthscales (D)]`` of ``K(A, B)`` from ``G = dL/dK`` in one fused, deterministic sweep
    (``cggp_kernel_matrix_backward``)."""
    G = _lib.row_major(G.contiguous())
    c = _lib.context(A.P.device)
    c.use_current_stream()
    ls = torch.as_tensor(lengthscales, dtype=torch.float64).reshape(-1).cpu()
    arr = (C.c_double * ls.numel())(*ls.tolist())
    g_var = torch.empty((1,), dtype=A.P.dtype, device=A.P.device)
    g_ls = torch.empty((A.D,), dtype=A.P.dtype, device=A.P.device)
    c.check(c.lib.cggp_kernel_matrix_backward(
        c.handle, _lib.dtype_code(A.P.dtype), int(kind), float(variance), _lib.ptr(A.P), A.n, _lib.ptr(B.P), B.n, A.D,
        A.P.shape[1], arr, ls.numel(), _lib.ptr(G), G.stride(0), _lib.ptr(g_var), _lib.ptr(g_ls)))
    return torch.cat([g_var, g_ls])


class _KernelMatrixFn(torch.autograd.Function):
    """``K(X, X2)`` differentiable in the hyper-parameters: forward = ``cggp_kernel_matrix``, backward =
    ``cggp_kernel_matrix_backward`` (dL/dvariance, dL/dlengthscales from dL/dK in one fused sweep).  No gradient is
    produced for the inputs: the reference trains kernel / likelihood parameters and moves the inducing points by
    clustering (cggp/optimize.py:19-98), not by gradient."""

    @staticmethod
    def forward(ctx, variance_t, lengthscales_t, X, X2, kind, jitter):
        ls = lengthscales_t.detach().reshape(-1).double().cpu()
        A = prepare_points(X, ls)
        B = A if X2 is None else prepare_points(X2, ls, A.P.dtype)
        var = float(variance_t.detach())
        ctx.save_for_backward(A.P, B.P)
        ctx.meta = (kind, var, ls, A.D, variance_t, lengthscales_t)
        return kernel_matrix(kind, var, A, B, jitter=jitter)

    @staticmethod
    def backward(ctx, G):
        PA, PB = ctx.saved_tensors
        kind, var, ls, D, variance_t, lengthscales_t = ctx.meta
        G = _lib.row_major(G.contiguous())
        c = _lib.context(PA.device)
        c.use_current_stream()
        g_var = torch.empty((1,), dtype=PA.dtype, device=PA.device)
        g_ls = torch.empty((D,), dtype=PA.dtype, device=PA.device)
        arr = (C.c_double * ls.numel())(*ls.tolist())
        c.check(c.lib.cggp_kernel_matrix_backward(
            c.handle, _lib.dtype_code(PA.dtype), int(kind), var, _lib.ptr(PA), PA.shape[0], _lib.ptr(PB), PB.shape[0],
            D, PA.shape[1], arr, ls.numel(), _lib.ptr(G), G.stride(0), _lib.ptr(g_var), _lib.ptr(g_ls)))
        if lengthscales_t.numel() == 1:
            g_ls = g_ls.sum()
        return (g_var.reshape(variance_t.shape).to(variance_t.dtype),
                g_ls.reshape(lengthscales_t.shape).to(lengthscales_t.dtype), None, None, None, None)


class _SGPRTermsFn(torch.autograd.Function):
    """``(K(Z, Z), Kuf Kfu, Kuf Y)`` over this rank's shard, differentiable in the kernel hyper-parameters: the three
    kernel-dependent terms of GPflow's ``SGPR.elbo``.  Forward forms ``Kuf`` in row chunks (``cggp_kernel_matrix``) and
    accumulates the rank-k updates; backward re-forms each chunk, builds ``dL/dKuf = (dG + dG^T) Kuf + dW Y^T`` and
    hands it to ``cggp_kernel_matrix_backward``.  The shard terms (forward values and parameter gradients) are
    all-reduced over the communicator of the context, so every rank sees the full gradient."""

    @staticmethod
    def forward(ctx, variance_t, lengthscales_t, X, Z, Y, kind):
        ls = lengthscales_t.detach().reshape(-1).double().cpu()
        var = float(variance_t.detach())
        PZ = prepare_points(Z, ls)
        PX = prepare_points(X, ls, PZ.P.dtype)
        M = PZ.n
        Kzz = kernel_matrix(kind, var, PZ, PZ)
        G = kuf_gram(kind, var, PZ, PX)  # symmetric rank-k updates on the library's DMMA GEMM
        W = torch.zeros((M, Y.shape[1]), dtype=PZ.P.dtype, device=PZ.P.device)
        step = max(1, (1 << 24) // max(M, 1))
        for s in range(0, PX.n, step):
            e = min(PX.n, s + step)
            W.addmm_(kernel_matrix(kind, var, PZ, PX.rows(s, e)), Y[s:e])  # [M, nc] @ [nc, P]: P columns, HBM-bound
        c = _lib.context(PZ.P.device)
        if c.world > 1:
            c.allreduce_sum_(G)
            c.allreduce_sum_(W)
        ctx.save_for_backward(PZ.P, PZ.norms, PX.P, PX.norms, Y)
        ctx.meta = (kind, var, ls, PZ.D, variance_t, lengthscales_t, step)
        return Kzz, G, W

    @staticmethod
    def backward(ctx, dKzz, dG, dW):
        PZp, PZn, PXp, PXn, Y = ctx.saved_tensors
        kind, var, ls, D, variance_t, lengthscales_t, step = ctx.meta
        c = _lib.context(PZp.device)
        c.use_current_stream()
        arr = (C.c_double * ls.numel())(*ls.tolist())
        PZ = PreparedPoints(PZp, PZn, D)
        PX = PreparedPoints(PXp, PXn, D)

        def kmb(PA, PB, Gm):
            Gm = _lib.row_major(Gm.contiguous())
            g_var = torch.empty((1,), dtype=PZp.dtype, device=PZp.device)
            g_ls = torch.empty((D,), dtype=PZp.dtype, device=PZp.device)
            c.check(c.lib.cggp_kernel_matrix_backward(
                c.handle, _lib.dtype_code(PZp.dtype), int(kind), var, _lib.ptr(PA.P), PA.n, _lib.ptr(PB.P), PB.n, D,
                PA.P.shape[1], arr, ls.numel(), _lib.ptr(Gm), Gm.stride(0), _lib.ptr(g_var), _lib.ptr(g_ls)))
            return torch.cat([g_var, g_ls])

        shard = torch.zeros((1 + D,), dtype=PZp.dtype, device=PZp.device)
        sym = None if dG is None else dG + dG.t()
        for s in range(0, PX.n, step):
            e = min(PX.n, s + step)
            rows = PX.rows(s, e)
            # dL/dKfu [nc, M] = Kfu (dG + dG^T) + Y dW^T: the N M^2 part on the library's DMMA GEMM (sym is symmetric)
            if sym is not None:
                dKt = symm_matmul(sym, kernel_matrix(kind, var, rows, PZ))
            else:
                dKt = torch.zeros((e - s, PZ.n), dtype=PZp.dtype, device=PZp.device)
            if dW is not None:
                dKt.addmm_(Y[s:e], dW.t())  # rank-P update
            shard += kmb(rows, PZ, dKt)
        if c.world > 1:
            c.allreduce_sum_(shard)
        g = shard
        if dKzz is not None:
            g = g + kmb(PZ, PZ, dKzz)  # replicated term: identical on every rank, not reduced
        g_var, g_ls = g[:1], g[1:]
        if lengthscales_t.numel() == 1:
            g_ls = g_ls.sum()
        return (g_var.reshape(variance_t.shape).to(variance_t.dtype),
                g_ls.reshape(lengthscales_t.shape).to(lengthscales_t.dtype), None, None, None, None)


class Stationary:
    """``variance`` / ``lengthscales`` may be numbers (fixed) or torch tensors that require grad (trainable: pass e.g.
    ``softplus(raw)``); with trainable parameters ``K`` / ``K_diag`` are differentiable (``_KernelMatrixFn``)."""

    kind = None
    name = "stationary"

    def __init__(self, variance=1.0, lengthscales=1.0):
        self._variance_t = variance if isinstance(variance, torch.Tensor) and variance.requires_grad else None
        self._lengthscales_t = lengthscales if isinstance(lengthscales, torch.Tensor) and lengthscales.requires_grad \
            else None
        self._variance = float(variance.detach()) if isinstance(variance, torch.Tensor) else float(variance)
        self._lengthscales = torch.as_tensor(lengthscales, dtype=torch.float64).detach().reshape(-1).cpu()

    @property
    def variance(self) -> float:
        return float(self._variance_t.detach()) if self._variance_t is not None else self._variance

    @property
    def lengthscales(self) -> torch.Tensor:
        if self._lengthscales_t is not None:
            return self._lengthscales_t.detach().reshape(-1).double().cpu()
        return self._lengthscales

    @property
    def trainable(self) -> bool:
        return self._variance_t is not None or self._lengthscales_t is not None

    @property
    def ard(self) -> bool:
        return self.lengthscales.numel() > 1

    def prepare(self, X, dtype=None) -> PreparedPoints:
        return X if isinstance(X, PreparedPoints) else prepare_points(X, self.lengthscales, dtype)

    def K(self, X, X2=None, *, jitter=0.0):
        if self.trainable and torch.is_grad_enabled() and not isinstance(X, PreparedPoints):
            Xt = _lib.as_device_tensor(X)
            X2t = None if X2 is None else _lib.as_device_tensor(X2, Xt.dtype)
            var_t = self._variance_t if self._variance_t is not None else \
                torch.tensor(self._variance, dtype=Xt.dtype, device=Xt.device)
            ls_t = self._lengthscales_t if self._lengthscales_t is not None else \
                self._lengthscales.to(Xt.device)
            return _KernelMatrixFn.apply(var_t, ls_t, Xt, X2t, self.kind, float(jitter))
        A = self.prepare(X)
        B = A if X2 is None else self.prepare(X2, A.P.dtype)
        return kernel_matrix(self.kind, self.variance, A, B, jitter=jitter)

    def param_tensors(self, dtype, device):
        """(variance, lengthscales) as tensors on `device` (the trainable ones where present)."""
        var_t = self._variance_t if self._variance_t is not None else \
            torch.tensor(self._variance, dtype=dtype, device=device)
        ls_t = self._lengthscales_t if self._lengthscales_t is not None else self._lengthscales.to(device)
        return var_t, ls_t

    def K_diag(self, X):
        if isinstance(X, PreparedPoints):
            n, dtype, device = X.n, X.P.dtype, X.P.device
        else:
            X = _lib.as_device_tensor(X)
            n, dtype, device = X.shape[0], X.dtype, X.device
        if self._variance_t is not None and torch.is_grad_enabled():
            return self._variance_t.to(device=device, dtype=dtype).reshape(1).expand(n)
        return torch.full((n,), self.variance, dtype=dtype, device=device)

    def __call__(self, X, X2=None, *, full_cov=True):
        if not full_cov:
            if X2 is not None:
                raise ValueError("Ambiguous inputs: `not full_cov` and `X2` are not compatible.")
            return self.K_diag(X)
        return self.K(X, X2)


class SquaredExponential(Stationary):
    kind = _lib.SE
    name = "se"


class Matern12(Stationary):
    kind = _lib.MATERN12
    name = "matern12"


class Matern32(Stationary):
    kind = _lib.MATERN32
    name = "matern32"


class Matern52(Stationary):
    kind = _lib.MATERN52
    name = "matern52"


KERNELS = {k.name: k for k in (SquaredExponential, Matern12, Matern32, Matern52)}


class InducingPoints:
    """GPflow ``InducingPoints`` surface: ``.Z`` and ``.num_inducing``."""

    def __init__(self, Z):
        self.Z = _lib.as_device_tensor(Z)

    @property
    def num_inducing(self) -> int:
        return self.Z.shape[0]


def inducingpoint_wrapper(iv) -> InducingPoints:
    return iv if isinstance(iv, InducingPoints) else InducingPoints(iv)


def Kuu(inducing_variable, kernel: Stationary, *, jitter=0.0):
    """GPflow ``covariances.Kuu``: ``K(Z, Z) + jitter I`` (cggp/models.py:300,333)."""
    return kernel.K(inducingpoint_wrapper(inducing_variable).Z, jitter=float(jitter))


def Kuf(inducing_variable, kernel: Stationary, Xnew):
    """GPflow ``covariances.Kuf``: ``K(Z, Xnew)`` [M, N] (cggp/models.py:334)."""
    return kernel.K(inducingpoint_wrapper(inducing_variable).Z, Xnew)


class Gaussian:
    """GPflow ``likelihoods.Gaussian`` (variance only), the likelihood of cggp/cli_utils.py:153,164."""

    def __init__(self, variance=1.0):
        # a number (fixed) or a torch scalar that requires grad (trainable noise variance)
        self.variance = variance if isinstance(variance, torch.Tensor) and variance.requires_grad else float(variance)

    def variational_expectations(self, X, Fmu, Fvar, Y):
        import math

        v = self.variance
        log_v = torch.log(v) if isinstance(v, torch.Tensor) else math.log(v)
        ve = -0.5 * math.log(2.0 * math.pi) - 0.5 * log_v - 0.5 * ((Y - Fmu) ** 2 + Fvar) / v
        return ve.sum(-1)

    def predict_log_density(self, X, Fmu, Fvar, Y):
        import math

        s2 = Fvar + self.variance
        ld = -0.5 * (math.log(2.0 * math.pi) + torch.log(s2) + (Y - Fmu) ** 2 / s2)
        return ld.sum(-1)
