"""The reference's known-answer test of the kernel values (``cggp/rff_test.py:9-29``, ``test_rff_kernel``): the
random-Fourier-feature estimate built by the reference's unmodified ``cggp/rff.py`` must reproduce ``K(X, X)``.  The
estimates in ``tests/golden/rff_golden.npz`` come from that file run here (tests/golden/make_golden_rff.py, 4e6 bases);
the restated GPflow kernels of the oracle are held to them at the reference test's own tolerance (rtol 1e-3, atol
1e-2) and at the Monte-Carlo level of the fixture (4e-3).  A kernel with a wrong scale constant fails the same check."""
import os

import numpy as np
import pytest

from oracle import gpflow_restated as g

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "rff_golden.npz"))
KEYS = sorted({"/".join(k.split("/")[:2]) for k in GOLD.files})


@pytest.mark.parametrize("key", KEYS)
def test_restated_kernels_match_the_reference_rff_estimates(key):
    name = key.split("/")[0]
    X, ls, var = GOLD[f"{key}/inputs"], GOLD[f"{key}/lengthscales"], float(GOLD[f"{key}/variance"])
    kxx = g.KERNELS[name](variance=var, lengthscales=ls).K(X)
    np.testing.assert_allclose(GOLD[f"{key}/rff_approx"], kxx, rtol=1e-3, atol=1e-2)   # rff_test.py:29
    assert np.abs(GOLD[f"{key}/rff_approx"] - kxx).max() < 4e-3                        # 4e6 bases: sigma ~ 7e-4


def test_the_rff_pin_tells_wrong_kernels_apart():
    key = "matern52/d2"
    X, ls, var = GOLD[f"{key}/inputs"], GOLD[f"{key}/lengthscales"], float(GOLD[f"{key}/variance"])
    rff = GOLD[f"{key}/rff_approx"]
    assert np.abs(g.Matern32(variance=var, lengthscales=ls).K(X) - rff).max() > 1e-2        # other smoothness
    assert np.abs(g.Matern52(variance=var, lengthscales=ls * 1.1).K(X) - rff).max() > 1e-2  # lengthscale convention
    assert np.abs(g.SquaredExponential(variance=var, lengthscales=ls).K(X) - rff).max() > 1e-2
