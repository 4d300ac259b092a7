"""Stand-in for the two tensorflow_probability symbols ``cggp/models.py`` imports (golden generation only)."""
import types

import numpy as np

distributions = types.SimpleNamespace()
_rng = np.random.default_rng(12345)


def _rademacher(shape, dtype=np.float64):
    return (2.0 * _rng.integers(0, 2, size=tuple(int(s) for s in shape)) - 1.0).astype(dtype)


random = types.SimpleNamespace(rademacher=_rademacher, _reseed=lambda s: globals().update(_rng=np.random.default_rng(s)))
