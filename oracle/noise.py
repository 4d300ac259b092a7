"""Rounding-noise floor of the reference algorithm itself, measured on the oracle (test infrastructure).

CG stopped at a finite threshold (or at max_iterations) returns a vector that depends on the ORDER in which the
floating-point sums of `p @ A` are formed: on the reference's own test systems (cond 3e3..1e5) two summation orders give
solutions differing by 1e-8 (threshold 1e-18) to 4e-3 (threshold 1e-6, the reference default).  The reference's TF path,
the NumPy oracle and the CUDA path are three such orders.  Parity is therefore asserted as
``|cuda - reference| <= base_tol + SLACK * noise`` with ``noise`` = the largest deviation between the oracle and the
same oracle with its matrix products summed in permuted orders."""
import numpy as np

from oracle import cg as ocg

SLACK = 20.0
SEEDS = (0, 1, 2, 3)


class PermutedCG(ocg.ConjugateGradient):
    """The oracle's ConjugateGradient with `V @ A` summed in a seeded permuted order (same algorithm, same maths)."""

    def __init__(self, seed, *args, **kw):
        super().__init__(*args, **kw)
        self.seed = seed

    def __call__(self, matrix, rhs, initial_solution=None):
        perm = np.random.default_rng(self.seed).permutation(matrix.shape[0])
        return super().__call__(lambda V: V[:, perm] @ matrix[perm, :], rhs, initial_solution)


def permuted_matmul(A, seed):
    perm = np.random.default_rng(seed).permutation(A.shape[0])
    return lambda V: V[:, perm] @ A[perm, :]


def deviation(ref, alts):
    """max_k max|alts[k] - ref| (0 for an empty list)."""
    return max([float(np.max(np.abs(np.asarray(a) - np.asarray(ref)))) for a in alts] + [0.0])


def assert_close_with_noise(actual, ref, alts, base_rtol, what=""):
    ref = np.asarray(ref)
    scale = float(np.max(np.abs(ref))) if ref.size else 0.0
    tol = base_rtol * scale + SLACK * deviation(ref, alts)
    err = float(np.max(np.abs(np.asarray(actual) - ref))) if ref.size else 0.0
    assert err <= tol, f"{what}: |cuda - reference| = {err:.3e} > {tol:.3e} (scale {scale:.3e}, noise {deviation(ref, alts):.3e})"


class PermutedDensePreconditioner(ocg.DensePreconditioner):
    """The oracle's DensePreconditioner with `vec @ Pinv` summed in a seeded permuted order."""

    def __init__(self, pinv, seed):
        super().__init__(pinv)
        self.perm = np.random.default_rng(1000 + seed).permutation(self.pinv.shape[0])

    def __call__(self, vec, mat):
        z = vec[:, self.perm] @ self.pinv[self.perm, :]
        return z, np.sum(z * vec, axis=-1, keepdims=True)
