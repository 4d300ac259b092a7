"""The reference's only known-answer test of the KERNEL VALUES, ``cggp/rff_test.py:9-29`` (``test_rff_kernel``): the
random-Fourier-feature estimate ``variance / L * phi phi^T`` built by the reference's UNMODIFIED ``cggp/rff.py``
(spectral sampling of SE / Matern kernels, independent of the closed-form kernel expressions) must reproduce ``K(X, X)``.

This script runs that construction over the NumPy stand-ins (``tests/golden/_shim``: ``tf.matmul / cos / sin``,
``tfd.MultivariateNormalDiag / Chi2`` on a seeded NumPy generator) with 4e6 bases instead of the test's 1e5 (Monte-Carlo
error ~1e-3 instead of ~4e-3) and stores inputs, kernel parameters and the estimates in ``tests/golden/rff_golden.npz``
(committed).  ``tests/test_oracle_rff.py`` holds the restated kernels (oracle/gpflow_restated.py) - and through them the
CUDA kernels - to those estimates at the reference test's own tolerance.  Run in the build container only."""
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
import rff as ref_rff  # noqa: E402  (reference, unmodified)
from gpflow.kernels import Matern12, Matern32, Matern52, SquaredExponential  # noqa: E402  (shim -> oracle classes)

assert ref_rff.__file__.startswith(REF)


def main():
    store = {}
    rng = np.random.default_rng(77)
    tfp.random._reseed(2024)
    num_bases = 4_000_000
    for name, cls in (("se", SquaredExponential), ("matern12", Matern12), ("matern32", Matern32), ("matern52", Matern52)):
        for dimension, num_inputs in ((2, 4), (3, 6)):         # rff_test.py:9 uses (2, 4)
            inputs = rng.standard_normal((num_inputs, dimension))
            lengthscales = rng.random(dimension) ** 2 + 0.5     # rff_test.py:18
            variance = 1.3                                      # rff_test.py:19
            kernel = cls(variance=variance, lengthscales=lengthscales)
            approx = np.zeros((num_inputs, num_inputs))
            for _ in range(8):                                  # in slabs: [n, 2 L] features of 5e5 bases at a time
                theta = ref_rff.basis_theta_parameter(kernel, num_bases=num_bases // 8)
                basis_vs = ref_rff.basis_vectors(inputs, theta=theta)
                approx += tf.matmul(basis_vs, basis_vs, transpose_b=True)
            approx *= tf.math.truediv(kernel.variance, num_bases)    # rff_test.py:24-25
            key = f"{name}/d{dimension}"
            store[f"{key}/inputs"] = inputs
            store[f"{key}/lengthscales"] = lengthscales
            store[f"{key}/variance"] = np.float64(variance)
            store[f"{key}/rff_approx"] = approx
            print(key, "max |rff - K| =", np.abs(approx - kernel(inputs)).max())
    np.savez_compressed(os.path.join(HERE, "rff_golden.npz"), **store)


if __name__ == "__main__":
    main()
