"""Mirror of ``cggp/models.py`` (objective + prediction) on the B200 path.

``CGGP`` ("CDGP", models.py:279-354) does every solve with the device CG; ``ClusterGP`` (models.py:176-276) is
the reference's exact Cholesky comparator (kept, as in the reference, for cross-checks: its factorisations are plain
``torch.linalg`` library calls, not part of the hot path); ``SGPR`` is the GPflow model the reference builds in
``cli_utils.sgpr_class`` (cli_utils.py:444-446), here solved matrix-free through the operator
``Kuu + s^-2 Kuf Kfu`` it factorises.  Entry points and shapes are the reference's:
``elbo(data)``, ``predict_f(Xnew, full_cov=False, full_output_cov=False) -> (mean [N,1], var [N,1] | [1,N,N])``,
``prior_kl()``, attributes ``pseudo_u``, ``cluster_counts``, ``diag_variance``, ``inducing_variable.Z``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .conjugate_gradient import ConjugateGradient, EyePreconditioner
from .gpflow_adapter import inducing_from_gpflow, kernel_from_gpflow, likelihood_from_gpflow
from .kernels import Gaussian, InducingPoints, Kuf, Kuu, Stationary, inducingpoint_wrapper
from .operators import SGPROperator
from .utils import add_diagonal

Tensor = torch.Tensor
Moments = Tuple[Tensor, Tensor]


def eval_logdet(matrix, cg: ConjugateGradient, num_probes=None, probes=None):
    """models.py:21-48: forward value is the constant 0; the gradient w.r.t. ``matrix`` is ``df * (A^-1)^T`` by CG
    against the identity, or the Hutchinson estimate with Rademacher probes."""

    class _LogDet(torch.autograd.Function):
        @staticmethod
        def forward(ctx, A):
            ctx.save_for_backward(A.detach())
            return torch.zeros((), dtype=A.dtype, device=A.device)  # :46

        @staticmethod
        def backward(ctx, df):
            (A,) = ctx.saved_tensors
            n = A.shape[-1]
            if num_probes is None:  # :32-36
                eye = torch.eye(n, dtype=A.dtype, device=A.device)
                return df * cg(A, eye).t()
            pr = probes if probes is not None else rademacher((n, num_probes), A.dtype, A.device)  # :38-39
            lv = cg(A, pr)
            return (lv @ (df * pr).t()) / num_probes  # :40-42

    return _LogDet.apply(_lib.as_device_tensor(matrix))


def rademacher(shape, dtype, device, generator=None):
    return (torch.randint(0, 2, shape, device=device, generator=generator) * 2 - 1).to(dtype)


class LpSVGP:
    """models.py:51-173: diagonal-parametrised SVGP (parent of ClusterGP / CGGP; provides ``elbo`` and ``scale``)."""

    def __init__(self, kernel: Stationary, likelihood: Gaussian, inducing_variable, *, mean_function=None,
                 num_latent_gps: int = 1, nu=None, diag_variance=None, num_data=None):
        assert num_latent_gps == 1, "One latent GP is allowed"
        # GPflow objects (gpflow.kernels.*, gpflow.likelihoods.Gaussian, InducingPoints) are accepted as in the
        # reference (models.py:279-291) and converted by duck typing; objects of this package pass through
        self.kernel = kernel_from_gpflow(kernel)
        self.likelihood = likelihood_from_gpflow(likelihood)
        self.mean_function = mean_function
        self.num_latent_gps = num_latent_gps
        self.num_data = num_data
        self.inducing_variable: InducingPoints = inducing_from_gpflow(inducing_variable)
        Z = self.inducing_variable.Z
        m = Z.shape[0]
        self._nu = torch.zeros((m, 1), dtype=Z.dtype, device=Z.device) if nu is None else \
            _lib.as_device_tensor(nu, Z.dtype).reshape(m, 1).clone()
        self._diag_variance = torch.full((m, 1), 1e-4, dtype=Z.dtype, device=Z.device) if diag_variance is None else \
            _lib.as_device_tensor(diag_variance, Z.dtype).reshape(m, 1).clone()

    @property
    def nu(self):
        return self._nu

    @property
    def diag_variance(self):
        return self._diag_variance

    def _mean(self, Xnew, like):
        if self.mean_function is None:
            return torch.zeros_like(like)
        return self.mean_function(Xnew)

    def prior_kl(self) -> Tensor:  # :107-120
        Kmm = Kuu(self.inducing_variable, self.kernel, jitter=0.0)
        nu, var = self.nu, self.diag_variance
        quad = (nu * (Kmm @ nu)).sum()
        K = add_diagonal(Kmm, var[:, 0])
        L = torch.linalg.cholesky(K)
        trace = torch.trace(torch.cholesky_solve(Kmm, L))
        logdet = (2.0 * torch.log(torch.diagonal(L))).sum() - torch.log(var).sum()
        return 0.5 * (quad - trace + logdet)

    def maximum_log_likelihood_objective(self, data) -> Tensor:  # :122-123
        return self.elbo(data)

    def elbo(self, data) -> Tensor:  # :125-134
        x, y = data
        x = _lib.as_device_tensor(x)
        y = _lib.as_device_tensor(y, x.dtype)
        kl = self.prior_kl()
        f_mean, f_var = self.predict_f(x, full_cov=False, full_output_cov=False)
        scale = self.scale(x.shape[0], kl.dtype)
        if not self._needs_grad() and y.shape[1] == 1:
            # forward value only: the Gaussian expectation sum as ONE deterministic device reduction (cggp_elbo_terms)
            ctx = _lib.context(x.device)
            ctx.use_current_stream()
            out = torch.empty((1,), dtype=x.dtype, device=x.device)
            yc, mc, vc = y.contiguous(), f_mean.contiguous(), f_var.contiguous()
            ctx.check(ctx.lib.cggp_elbo_terms(ctx.handle, _lib.dtype_code(x.dtype), _lib.ptr(yc), _lib.ptr(mc),
                                              _lib.ptr(vc), yc.shape[0], float(self.likelihood.variance), _lib.ptr(out)))
            return out[0] * scale - kl
        var_exp = self.likelihood.variational_expectations(x, f_mean, f_var, y)
        return var_exp.sum() * scale - kl

    def _needs_grad(self) -> bool:
        """True when a differentiable evaluation is wanted (trainable kernel / likelihood parameters under grad mode):
        the fused forward-only C entry points (cggp_predict_f, cggp_elbo_terms) are used otherwise."""
        if not torch.is_grad_enabled():
            return False
        v = self.likelihood.variance
        return bool(self.kernel.trainable or (isinstance(v, torch.Tensor) and v.requires_grad))

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False) -> Moments:  # :136-161
        assert not full_output_cov
        Xnew = _lib.as_device_tensor(Xnew)
        Kmm = Kuu(self.inducing_variable, self.kernel, jitter=0.0)
        Kmn = Kuf(self.inducing_variable, self.kernel, Xnew)
        Knn = self.kernel.K(Xnew) if full_cov else self.kernel.K_diag(Xnew)
        K = add_diagonal(Kmm, self.diag_variance[:, 0])
        L = torch.linalg.cholesky(K)
        A = torch.linalg.solve_triangular(L, Kmn, upper=False)
        if not full_cov:
            fvar = (Knn - (A * A).sum(0))[:, None]
        else:
            fvar = (Knn - A.t() @ A)[None, ...]
        fmu = Kmn.t() @ self.nu
        return fmu + self._mean(Xnew, fmu), fvar

    def scale(self, batch_size, dtype):  # :163-169
        if self.num_data is not None:
            return torch.tensor(float(self.num_data), dtype=dtype) / float(batch_size)
        return torch.tensor(1.0, dtype=dtype)

    def q_moments(self, full_cov: bool = False) -> Moments:  # :171-173
        return self.predict_f(self.inducing_variable.Z, full_cov=full_cov)


class ClusterGP(LpSVGP):
    """models.py:176-276: pseudo-targets = cluster means, ``Lambda = s2 / counts``; Cholesky solves."""

    def __init__(self, kernel, likelihood, inducing_variable, *, mean_function=None, num_latent_gps: int = 1,
                 cluster_counts=None, num_data=None, pseudo_u=None):
        assert num_latent_gps == 1, "One latent GP is allowed"
        super().__init__(kernel, likelihood, inducing_variable, num_latent_gps=num_latent_gps,
                         mean_function=mean_function, num_data=num_data)
        self.pseudo_u = self._nu  # :198
        del self._nu
        del self._diag_variance
        if pseudo_u is not None:  # :203-206
            pseudo_u = _lib.as_device_tensor(pseudo_u, self.pseudo_u.dtype)
            if tuple(pseudo_u.shape) != tuple(self.pseudo_u.shape):
                raise ValueError("Pseudo-u argument shape must match actual pseudo-u shape.")
            self.pseudo_u.copy_(pseudo_u)
        if cluster_counts is not None:  # :208-211
            cluster_counts = _lib.as_device_tensor(cluster_counts, self.pseudo_u.dtype)
            if tuple(cluster_counts.shape) != tuple(self.pseudo_u.shape):
                raise ValueError("Cluster counts argument shape must match pseudo-u shape.")
            counts = cluster_counts.clone()
        else:
            counts = torch.ones_like(self.pseudo_u)  # :213
        self.cluster_counts = counts

    @property
    def nu(self):  # :222-224
        raise NotImplementedError(f"This property is not supported in {self.__class__}")

    @property
    def diag_variance(self) -> Tensor:  # :226-228
        return self.likelihood.variance / self.cluster_counts

    def prior_kl(self) -> Tensor:  # :230-248
        Kmm = Kuu(self.inducing_variable, self.kernel, jitter=0.0)
        var = self.diag_variance
        K = add_diagonal(Kmm, var[:, 0])
        L = torch.linalg.cholesky(K)
        a = torch.cholesky_solve(self.pseudo_u, L)
        quad = ((Kmm @ a) * a).sum()
        trace = torch.trace(torch.cholesky_solve(Kmm, L))
        logdet = (2.0 * torch.log(torch.diagonal(L))).sum()
        const = torch.log(var).sum()
        return 0.5 * (quad - trace + logdet - const)

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False) -> Moments:  # :250-276
        assert not full_output_cov
        Xnew = _lib.as_device_tensor(Xnew)
        Kmm = Kuu(self.inducing_variable, self.kernel, jitter=0.0)
        Kmn = Kuf(self.inducing_variable, self.kernel, Xnew)
        Knn = self.kernel.K(Xnew) if full_cov else self.kernel.K_diag(Xnew)
        K = add_diagonal(Kmm, self.diag_variance[:, 0])
        L = torch.linalg.cholesky(K)
        a = torch.cholesky_solve(self.pseudo_u, L)
        A = torch.linalg.solve_triangular(L, Kmn, upper=False)
        if not full_cov:
            fvar = (Knn - (A * A).sum(0))[:, None]
        else:
            fvar = (Knn - A.t() @ A)[None, ...]
        fmu = Kmn.t() @ a
        return fmu + self._mean(Xnew, fmu), fvar


class CGGP(ClusterGP):
    """models.py:279-354 (the CLI's "cdgp"): ``A = Kuu + Lambda`` solved by CG for ``pseudo_u`` (1 RHS), the probes /
    ``Kmm`` (trace term) and ``Kmn`` (predictive variance, B RHS: the hot ``[B, M] @ [M, M]`` loop)."""

    def __init__(self, kernel, likelihood, inducing_variable, conjugate_gradient: ConjugateGradient,
                 num_probes: Optional[int] = 5, **kwargs):
        super().__init__(kernel, likelihood, inducing_variable, **kwargs)
        self.conjugate_gradient = conjugate_gradient
        self.num_probes = num_probes
        self.probes = None            # optional injected Rademacher probes [M, P] (the reference uses TF's global RNG)
        self.probe_generator = None   # optional torch.Generator for reproducible probes

    def _draw_probes(self, n, dtype, device):
        if self.probes is not None:
            return _lib.as_device_tensor(self.probes, dtype)
        return rademacher((n, self.num_probes), dtype, device, self.probe_generator)

    def prior_kl(self) -> Tensor:  # :293-322
        pseudo_u = self.pseudo_u
        var = self.diag_variance
        Kmm = Kuu(self.inducing_variable, self.kernel, jitter=0.0)  # :300
        KmmLambda = add_diagonal(Kmm, var[:, 0])  # :301
        cg = self.conjugate_gradient
        a = cg(KmmLambda, pseudo_u)  # :303
        if self.num_probes is None:  # :304-306
            trace = torch.trace(cg(KmmLambda, Kmm))
        else:  # :308-314
            probes = self._draw_probes(KmmLambda.shape[0], KmmLambda.dtype, KmmLambda.device)
            sol = cg(KmmLambda, probes)
            trace = (sol * (Kmm @ probes)).sum() / self.num_probes
        quad = ((Kmm @ a) * a).sum()  # :316-317
        logdet = eval_logdet(KmmLambda, cg, num_probes=self.num_probes)  # :319 (value 0)
        const = torch.log(var).sum()  # :321
        return 0.5 * (quad - trace + logdet - const)

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False) -> Moments:  # :324-354
        assert not full_output_cov
        Xnew = _lib.as_device_tensor(Xnew)
        var = self.diag_variance
        Kmm = Kuu(self.inducing_variable, self.kernel, jitter=0.0)  # :333
        Knn = self.kernel.K(Xnew) if full_cov else self.kernel.K_diag(Xnew)  # :335
        KmmLambda = add_diagonal(Kmm, var[:, 0])  # :337
        cg = self.conjugate_gradient
        a = cg(KmmLambda, self.pseudo_u)  # :339
        if not full_cov and not self._needs_grad() and isinstance(cg.preconditioner, EyePreconditioner) \
                and not cg.record_history and self.mean_function is None and Xnew.shape[0] > 0:
            return self._predict_f_fused(Xnew, KmmLambda, a)
        Kmn = Kuf(self.inducing_variable, self.kernel, Xnew)  # :334
        S = cg(KmmLambda, Kmn)  # :340
        if not full_cov:
            fvar = (Knn - (Kmn * S).sum(0))[:, None]  # :343-345
        else:
            fvar = (Knn - Kmn.t() @ S)[None, ...]  # :347-349
        fmu = Kmn.t() @ a  # :351
        return fmu + self._mean(Xnew, fmu), fvar


    def _predict_f_fused(self, Xnew, KmmLambda, a) -> Moments:
        """models.py:334-352 as ONE call of the C ABI (``cggp_predict_f``): Kmn, the B-RHS solve with the reference's
        stopping rule, ``fvar = Knn - sum_m Kmn * S`` and ``fmu = Kmn^T a`` - forward values, nothing leaves the
        device.  The differentiable path above is kept for training."""
        import ctypes as C

        cg = self.conjugate_gradient
        Z = self.inducing_variable.Z
        PZ = self.kernel.prepare(Z)
        PN = self.kernel.prepare(Xnew, PZ.P.dtype)
        ctx = _lib.context(Z.device)
        ctx.use_current_stream()
        nb, m = PN.n, PZ.n
        Knm = torch.empty((nb, m), dtype=PZ.P.dtype, device=Z.device)
        S = torch.empty_like(Knm)
        mean = torch.empty((nb,), dtype=PZ.P.dtype, device=Z.device)
        var = torch.empty_like(mean)
        max_it = cg.max_iterations if cg.max_iterations is not None else m
        cycle = cg.max_steps_cycle if cg.max_steps_cycle is not None else max_it + 1
        A = _lib.row_major(KmmLambda)
        av = a.reshape(-1).contiguous()
        steps = C.c_int32(0)
        ctx.check(ctx.lib.cggp_predict_f(
            ctx.handle, _lib.dtype_code(PZ.P.dtype), self.kernel.kind, self.kernel.variance, _lib.ptr(PZ.P),
            _lib.ptr(PZ.norms), m, _lib.ptr(PN.P), _lib.ptr(PN.norms), nb, PZ.D, PZ.ldp, _lib.ptr(A), A.stride(0),
            _lib.ptr(av), float(cg.error_threshold), int(max_it), int(cycle), _lib.ptr(Knm), _lib.ptr(S), _lib.ptr(mean),
            _lib.ptr(var), C.byref(steps)))
        self.last_predict_steps = int(steps.value)
        return mean[:, None], var[:, None]


class SGPR:
    """GPflow ``models.SGPR`` as built by ``cli_utils.sgpr_class`` (cli_utils.py:444-446), zero mean function,
    evaluated through the linear system GPflow factorises, ``Sigma = Kuu + jitter I + s^-2 Kuf Kfu = L B L^T``:

        mean = Ksu Sigma^-1 Kuf y / s2            var = k** - Ksu Kuu^-1 Kus + Ksu Sigma^-1 Kus

    ``Sigma`` is applied matrix-free (``SGPROperator``): Kfu [N, M] is never materialised (the reference needs all of
    it, 1 TB at the N = 8M / M = 16384 config).  ``data`` is THIS RANK's shard of (X, Y)."""

    def __init__(self, data, kernel: Stationary, inducing_variable, *, noise_variance: float = 1.0,
                 conjugate_gradient: Optional[ConjugateGradient] = None, jitter: float = 1e-6, variant: int = 0):
        X, Y = data
        self.X = _lib.as_device_tensor(X)
        self.Y = _lib.as_device_tensor(Y, self.X.dtype)
        self.kernel = kernel_from_gpflow(kernel)
        kernel = self.kernel
        self.inducing_variable = inducing_from_gpflow(inducing_variable, self.X.dtype)
        self.likelihood = Gaussian(noise_variance)
        self.jitter = jitter
        self.conjugate_gradient = conjugate_gradient or ConjugateGradient(1e-6)
        nv = float(noise_variance.detach()) if isinstance(noise_variance, torch.Tensor) else float(noise_variance)
        self.operator = SGPROperator(kernel, self.X, self.inducing_variable.Z, nv, jitter, variant)
        self._c = None
        # predict_f solves Sigma S = Kus with one right-hand side per test point.  Matrix-free, every PAIR of right-hand
        # sides costs a full sweep over the N x M Gram entries per iteration; from `dense_threshold` test points on the
        # [M, M] matrix Sigma = Kuu + Kuf Kfu / s2 is formed once (N M^2 flop, all-reduced, cached) and the multi-RHS CG
        # runs on it with the DMMA GEMM - GPflow's SGPR materialises the same matrix (as L B L^T).
        self.dense_threshold = 16
        self._sigma_dense = None

    def _sync(self):
        """The operator and the cached solves are valid for ONE value of (kernel parameters, likelihood variance):
        after an in-place optimiser step they are re-prepared / dropped here instead of being reused silently."""
        s2 = self.likelihood.variance
        nv = float(s2.detach()) if isinstance(s2, torch.Tensor) else float(s2)
        if self.operator.stale or nv != self.operator.noise_variance:
            self.operator.refresh(s2)
            self._c = None
            self._sigma_dense = None

    def _posterior_weights(self):
        self._sync()
        if self._c is None:
            rhs = self.operator.kuf_times(self.Y) / self.likelihood.variance  # s^-2 Kuf y  [M, P]
            self._c = self.conjugate_gradient(self.operator, rhs)
        return self._c

    def elbo(self) -> Tensor:
        """GPflow ``SGPR.elbo`` (Titsias' collapsed bound; the objective behind ``cli_utils.sgpr_class``,
        cli_utils.py:444-446), same terms in the same order as GPflow: needs ``log det B`` and ``tr(A A^T)``, hence the
        materialised ``[M, M]`` Gram matrix ``Kuf Kfu`` (all-reduced over ranks) and two ``M x M`` Cholesky
        factorisations (library calls; this is the reference's own algorithm for this quantity)."""
        import math

        self._sync()
        op, ctx = self.operator, _lib.context(self.operator.device)
        s2 = self.likelihood.variance
        P = self.Y.shape[1]
        trainable = torch.is_grad_enabled() and (self.kernel.trainable or isinstance(s2, torch.Tensor))
        if trainable:
            # differentiable in (variance, lengthscales, noise variance): the kernel terms come from _SGPRTermsFn with
            # the CURRENT hyper-parameters (self.operator was prepared with the values at construction time)
            from .kernels import _SGPRTermsFn

            var_t, ls_t = self.kernel.param_tensors(self.X.dtype, self.X.device)
            Kzz, G, KufY = _SGPRTermsFn.apply(var_t, ls_t, self.X, self.inducing_variable.Z, self.Y, self.kernel.kind)
            Kuu_j = Kzz + self.jitter * torch.eye(op.n, dtype=Kzz.dtype, device=Kzz.device)
            kvar = var_t
            log_s2 = torch.log(s2) if isinstance(s2, torch.Tensor) else math.log(s2)
        else:
            G = op.gram()
            KufY = op.kuf_times(self.Y)  # [M, P], all-reduced
            Kuu_j, kvar, log_s2 = op.Kuu, self.kernel.variance, math.log(s2)
        stats = torch.stack([torch.tensor(float(self.X.shape[0]), dtype=self.Y.dtype, device=self.Y.device),
                             (self.Y * self.Y).sum()])
        if ctx.world > 1:
            ctx.allreduce_sum_(stats)
        n_total, sum_y2 = float(stats[0]), stats[1]
        L = torch.linalg.cholesky(Kuu_j)
        AAT = torch.linalg.solve_triangular(L, torch.linalg.solve_triangular(L, G, upper=False).t(), upper=False) / s2
        AAT = 0.5 * (AAT + AAT.t())
        B = AAT + torch.eye(op.n, dtype=AAT.dtype, device=AAT.device)
        LB = torch.linalg.cholesky(B)
        Aerr = torch.linalg.solve_triangular(L, KufY, upper=False) / s2  # A @ (Y / sigma)
        c = torch.linalg.solve_triangular(LB, Aerr, upper=False)
        const = -0.5 * n_total * P * math.log(2.0 * math.pi)
        half_logdet_B = torch.log(torch.diagonal(LB)).sum()
        trace_k = n_total * kvar / s2
        trace_q = torch.trace(AAT)
        logdet = -P * (half_logdet_B + 0.5 * n_total * log_s2 + 0.5 * (trace_k - trace_q))
        quad = -0.5 * (sum_y2 / s2 - (c * c).sum())
        return const + logdet + quad

    def maximum_log_likelihood_objective(self) -> Tensor:
        return self.elbo()

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False) -> Moments:
        assert not full_output_cov and not full_cov
        self._sync()
        Xnew = _lib.as_device_tensor(Xnew, self.X.dtype)
        Kus = Kuf(self.inducing_variable, self.kernel, Xnew)  # [M, B]
        c = self._posterior_weights()
        mean = Kus.t() @ c
        cg = self.conjugate_gradient
        S1 = cg(self.operator.Kuu, Kus)
        if self.dense_threshold is not None and Kus.shape[1] >= self.dense_threshold:
            if self._sigma_dense is None:
                self._sigma_dense = self.operator.Kuu + self.operator.gram() / self.likelihood.variance
            S2 = cg(self._sigma_dense, Kus)
        else:
            S2 = cg(self.operator, Kus)
        var = self.kernel.K_diag(Xnew) - (Kus * S1).sum(0) + (Kus * S2).sum(0)
        return mean, var[:, None].repeat(1, self.Y.shape[1])
