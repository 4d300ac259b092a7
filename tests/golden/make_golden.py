"""Generate golden vectors by running the reference's UNMODIFIED ``cggp/conjugate_gradient.py`` and
``cggp/models.py`` (from /root/reference, read-only) over the NumPy shim in ``tests/golden/_shim``.

Run in the build container only:  ``python tests/golden/make_golden.py``  ->  ``tests/golden/*.npz`` (committed).
Nothing in tests/, smoke() or bench.py reads /root/reference at run time; they read the .npz files.

What is pinned: the reference's own CG loop / guards / reset rule / stats / layout adapter / custom-gradient
closure, and its CGGP + ClusterGP ``prior_kl`` / ``predict_f`` / ``elbo`` / ``eval_logdet`` formulas.
What is NOT pinned: GPflow kernel arithmetic (supplied to the shim by ``oracle/gpflow_restated.py``).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/cggp"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, REF)

import tensorflow as tf  # noqa: E402  (the shim)
import tensorflow_probability as tfp  # noqa: E402  (the shim)
import conjugate_gradient as ref_cg  # noqa: E402  (reference, unmodified)
import models as ref_models  # noqa: E402  (reference, unmodified)
import gpflow  # noqa: E402  (the shim)

from oracle import gpflow_restated as g  # noqa: E402

assert ref_cg.__file__.startswith(REF) and ref_models.__file__.startswith(REF)


def history_from_trace():
    """0.5*|r|^2 per RHS at every evaluation of the reference's stopping condition."""
    hist = [0.5 * np.sum(np.square(lv[0].r), axis=-1) for lv in tf.while_loop_trace]
    tf.while_loop_trace.clear()
    return np.array(hist)


def spd_case(seed, n, d, kernel_name, dtype, diag):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(dtype)
    ls = (rng.random(d) ** 2 + 0.5).astype(dtype)
    k = g.KERNELS[kernel_name](variance=1.3, lengthscales=ls, dtype=dtype)
    A = k.K(X)
    A[np.arange(n), np.arange(n)] += np.asarray(diag, dtype=dtype)
    return rng, A.astype(dtype)


def gen_cg():
    out = {}
    cases = [
        # name, seed, n, m, d, kernel, dtype, diag, thr, max_it, cycle, x0, zero_row
        ("cgtest_se", 0, 100, 5, 2, "se", np.float64, 0.1 ** 2, 1e-12, None, 100, False, False),
        ("matern32_thr1e-6", 1, 128, 3, 3, "matern32", np.float64, 0.1, 1e-6, None, 129, False, False),
        ("reset_cycle7", 2, 96, 4, 2, "matern52", np.float64, 0.05, 1e-10, 60, 7, True, False),
        ("maxit_cap", 3, 80, 2, 2, "se", np.float64, 1e-3, 1e-14, 25, 100, False, False),
        ("zero_rhs_row", 4, 64, 3, 2, "matern12", np.float64, 0.1, 1e-8, None, 100, False, True),
        ("float32", 5, 96, 4, 3, "matern32", np.float32, 0.1, 1e-5, None, 100, False, False),
    ]
    for (name, seed, n, m, d, kern, dtype, diag, thr, max_it, cycle, use_x0, zero_row) in cases:
        rng, A = spd_case(seed, n, d, kern, dtype, diag)
        rhs = rng.standard_normal((m, n)).astype(dtype)
        if zero_row:
            rhs[1] = 0.0
        x0 = (0.1 * rng.standard_normal((m, n))).astype(dtype) if use_x0 else np.zeros_like(rhs)
        tf.while_loop_trace.clear()
        tf.last_custom_gradient.clear()
        sol, (steps, err) = ref_cg.conjugate_gradient(A, rhs, x0, thr, None, max_it, cycle)
        hist = history_from_trace()
        # backward closure (conjugate_gradient.py:100-118) on a seeded upstream sensitivity
        dx = rng.standard_normal((m, n)).astype(dtype)
        dA, db, _ = tf.last_custom_gradient[-1](dx, None, None)
        tf.while_loop_trace.clear()
        out.update({
            f"{name}/A": A, f"{name}/rhs": rhs, f"{name}/x0": x0,
            f"{name}/thr": np.float64(thr), f"{name}/max_it": np.int64(-1 if max_it is None else max_it),
            f"{name}/cycle": np.int64(cycle),
            f"{name}/solution": sol, f"{name}/steps": np.int32(steps), f"{name}/error": err,
            f"{name}/history": hist, f"{name}/dx": dx, f"{name}/dA": dA, f"{name}/db": db,
        })
        print(f"cg {name}: steps={int(steps)} max_err={float(np.max(err)):.3e} hist={hist.shape}")
    # ConjugateGradient adapter (column layout, never-refresh default)
    rng, A = spd_case(7, 72, 2, "matern32", np.float64, 0.1)
    rhs_cols = rng.standard_normal((72, 6))
    tf.while_loop_trace.clear()
    sol = ref_cg.ConjugateGradient(1e-6)(A, rhs_cols)
    hist = history_from_trace()
    out.update({"adapter/A": A, "adapter/rhs": rhs_cols, "adapter/thr": np.float64(1e-6),
                "adapter/solution": sol, "adapter/history": hist})
    print(f"cg adapter: iters={hist.shape[0] - 1}")
    np.savez_compressed(os.path.join(HERE, "cg_golden.npz"), **out)


def make_problem(seed, n, m, d, kernel_name, dtype=np.float64):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, size=(n, d)).astype(dtype)
    y = (np.sin(X.sum(-1, keepdims=True)) + 0.3 * rng.standard_normal((n, 1))).astype(dtype)
    Z = X[rng.choice(n, m, replace=False)].copy()
    d2 = g.square_distance(Z, X)
    idx = np.argmin(d2, axis=0)
    counts = np.maximum(np.bincount(idx, minlength=m), 1).astype(dtype)[:, None]
    u = np.zeros((m, 1), dtype)
    np.add.at(u, (idx, 0), y[:, 0])
    u = u / counts
    ls = (0.7 + 0.6 * rng.random(d)).astype(dtype)
    kernel = g.KERNELS[kernel_name](variance=1.2, lengthscales=ls, dtype=dtype)
    Xnew = rng.uniform(-3, 3, size=(37, d)).astype(dtype)
    return X, y, Z, u, counts, kernel, ls, Xnew


def gen_models():
    out = {}
    for name, seed, n, m, d, kern, thr, probes in [
        ("se_exacttrace", 11, 600, 48, 2, "se", 1e-10, None),
        ("matern32_probes", 12, 500, 40, 3, "matern32", 1e-6, 5),
        ("matern52_exacttrace", 13, 400, 56, 2, "matern52", 1e-8, None),
    ]:
        X, y, Z, u, counts, kernel, ls, Xnew = make_problem(seed, n, m, d, kern)
        lik = gpflow.likelihoods.Gaussian(variance=0.1)
        cg = ref_cg.ConjugateGradient(thr)
        tfp.random._reseed(1000 + seed)
        model = ref_models.CGGP(kernel, lik, Z.copy(), cg, num_probes=probes, num_data=n)
        model.pseudo_u.assign(u)
        model.cluster_counts.assign(counts)
        cl = ref_models.ClusterGP(kernel, lik, Z.copy(), num_data=n)
        cl.pseudo_u.assign(u)
        cl.cluster_counts.assign(counts)
        # the probes CGGP.prior_kl will draw (global RNG in the reference, models.py:310): replay the stream
        used_probes = np.zeros((m, 0))
        if probes is not None:
            tfp.random._reseed(1000 + seed)
            used_probes = tfp.random.rademacher((m, probes), dtype=np.float64)
            tfp.random._reseed(1000 + seed)
        kl = model.prior_kl()
        mu, var = model.predict_f(Xnew)
        mu_fc, var_fc = model.predict_f(Xnew, full_cov=True)
        tfp.random._reseed(1000 + seed)
        batch = (X[:200], y[:200])
        elbo = model.elbo(batch)
        ckl = cl.prior_kl()
        cmu, cvar = cl.predict_f(Xnew)
        celbo = cl.elbo(batch)
        # eval_logdet (models.py:21-48): forward value + exact-solve gradient closure
        Kmm = g.Kuu(Z, kernel)
        KmmLambda = Kmm.copy()
        KmmLambda[np.arange(m), np.arange(m)] += (0.1 / counts)[:, 0]
        tf.last_custom_gradient.clear()
        ld = ref_models.eval_logdet(KmmLambda, cg, num_probes=None)
        ld_grad = tf.last_custom_gradient[-1](np.float64(1.0))
        out.update({
            f"{name}/X": X, f"{name}/y": y, f"{name}/Z": Z, f"{name}/u": u, f"{name}/counts": counts,
            f"{name}/lengthscales": ls, f"{name}/variance": np.float64(1.2), f"{name}/noise": np.float64(0.1),
            f"{name}/Xnew": Xnew, f"{name}/thr": np.float64(thr), f"{name}/probes": used_probes,
            f"{name}/kernel": np.array(kern), f"{name}/num_data": np.int64(n),
            f"{name}/cggp_kl": np.float64(kl), f"{name}/cggp_mu": mu, f"{name}/cggp_var": var,
            f"{name}/cggp_var_fullcov": var_fc, f"{name}/cggp_elbo": np.float64(elbo),
            f"{name}/cluster_kl": np.float64(ckl), f"{name}/cluster_mu": cmu, f"{name}/cluster_var": cvar,
            f"{name}/cluster_elbo": np.float64(celbo),
            f"{name}/logdet_value": np.float64(ld), f"{name}/logdet_grad": ld_grad,
        })
        print(f"model {name}: kl={kl:.6f} elbo={elbo:.6f} cluster_kl={ckl:.6f} cluster_elbo={celbo:.6f}")
    tf.while_loop_trace.clear()
    np.savez_compressed(os.path.join(HERE, "models_golden.npz"), **out)


if __name__ == "__main__":
    gen_cg()
    gen_models()
