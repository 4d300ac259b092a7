"""Stand-in for the tensorflow_probability symbols ``cggp/models.py`` (rademacher) and ``cggp/rff.py``
(MultivariateNormalDiag, Chi2) use (golden generation only)."""
import types

import numpy as np

_rng = np.random.default_rng(12345)


class _MultivariateNormalDiag:
    def __init__(self, scale_diag):
        self.scale = np.asarray(scale_diag)

    def sample(self, sample_shape):
        return _rng.standard_normal(tuple(sample_shape) + self.scale.shape).astype(self.scale.dtype) * self.scale


class _Chi2:
    def __init__(self, df):
        self.df = np.asarray(df)

    def sample(self, sample_shape):
        return _rng.chisquare(float(self.df), size=tuple(sample_shape)).astype(self.df.dtype)


distributions = types.SimpleNamespace(MultivariateNormalDiag=_MultivariateNormalDiag, Chi2=_Chi2)


def _rademacher(shape, dtype=np.float64):
    return (2.0 * _rng.integers(0, 2, size=tuple(int(s) for s in shape)) - 1.0).astype(dtype)


random = types.SimpleNamespace(rademacher=_rademacher, _reseed=lambda s: globals().update(_rng=np.random.default_rng(s)))
