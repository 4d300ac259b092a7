"""Stand-in for the GPflow symbols ``cggp/models.py`` / ``cggp/rff.py`` touch (golden generation only).

The kernel / likelihood ARITHMETIC is ``oracle/gpflow_restated.py`` (GPflow itself is not installable here, so
that boundary stays unpinned); this file only provides the class plumbing (GPModel, Parameter, Kuu/Kuf
dispatch) that lets the reference's model code run unmodified.
"""
import types

import numpy as np

from oracle import gpflow_restated as _g

_float = [np.float64]


class Parameter(np.ndarray):
    def __new__(cls, value, dtype=None, shape=None, transform=None, **kw):
        return np.array(value, dtype=dtype).view(cls)

    def assign(self, value):
        self[...] = np.asarray(value, dtype=self.dtype)
        return self


class _InducingPoints:
    def __init__(self, Z):
        self.Z = Z if isinstance(Z, Parameter) else Parameter(Z)

    @property
    def num_inducing(self):
        return self.Z.shape[0]


class _GPModel:
    def __init__(self, kernel, likelihood, mean_function=None, num_latent_gps=1):
        self.kernel = kernel
        self.likelihood = likelihood
        self.mean_function = (lambda X: np.zeros((np.shape(X)[0], 1), dtype=np.asarray(X).dtype)) \
            if mean_function is None else mean_function
        self.num_latent_gps = num_latent_gps


class _Mixin:
    pass


kernels = types.SimpleNamespace(
    Kernel=_g.Stationary, SquaredExponential=_g.SquaredExponential, Matern12=_g.Matern12,
    Matern32=_g.Matern32, Matern52=_g.Matern52)
likelihoods = types.SimpleNamespace(Gaussian=_g.Gaussian)
covariances = types.SimpleNamespace(
    Kuu=lambda iv, kernel, jitter=0.0: _g.Kuu(np.asarray(iv.Z), kernel, jitter=float(jitter)),
    Kuf=lambda iv, kernel, Xnew: _g.Kuf(np.asarray(iv.Z), kernel, np.asarray(Xnew)))
models = types.SimpleNamespace(
    GPModel=_GPModel, ExternalDataTrainingLossMixin=_Mixin,
    util=types.SimpleNamespace(inducingpoint_wrapper=lambda z: z if isinstance(z, _InducingPoints) else _InducingPoints(z)))
base = types.SimpleNamespace(RegressionData=tuple)
utilities = types.SimpleNamespace(positive=lambda: None, set_trainable=lambda *a, **k: None)
config = types.SimpleNamespace(default_float=lambda: _float[0])
