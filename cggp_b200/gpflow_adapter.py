"""Adapter from GPflow objects to the objects of this package (the reference passes GPflow kernels,
``InducingPoints`` and likelihoods into its models, cggp/models.py:279-291,300,333-335; cggp/cli_utils.py:363-368).

GPflow is NOT imported: objects are recognised by duck typing (class name + attributes), so the adapter works with
real ``gpflow.kernels.*`` / ``gpflow.inducing_variables.InducingPoints`` / ``gpflow.likelihoods.Gaussian`` instances
(whose parameters are ``tf.Variable``-backed ``gpflow.Parameter`` objects exposing ``.numpy()`` and ``__dlpack__``)
as well as with any stand-in offering the same surface.  Parameter VALUES are read once (a snapshot, as
``SGPROperator`` does); device tensors (``Z``) are imported zero-copy through DLPack where the object offers it.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .kernels import Gaussian, InducingPoints, Matern12, Matern32, Matern52, SquaredExponential, Stationary

_KERNEL_BY_NAME = {
    "SquaredExponential": SquaredExponential, "RBF": SquaredExponential, "Exponential": Matern12,
    "Matern12": Matern12, "Matern32": Matern32, "Matern52": Matern52,
}


def _to_numpy(value) -> np.ndarray:
    """float64 host copy of a hyper-parameter: gpflow.Parameter / tf.Variable (``.numpy()``), torch, numpy, scalars."""
    if isinstance(value, torch.Tensor):
        return value.detach().double().cpu().numpy()
    if hasattr(value, "numpy") and callable(value.numpy):
        value = value.numpy()
    return np.asarray(value, dtype=np.float64)


def kernel_from_gpflow(kernel) -> Stationary:
    if isinstance(kernel, Stationary):
        return kernel
    name = type(kernel).__name__
    if name not in _KERNEL_BY_NAME or not hasattr(kernel, "variance") or not hasattr(kernel, "lengthscales"):
        raise TypeError(f"unsupported kernel {name!r}: the B200 path implements the stationary kernels the reference "
                        f"instantiates ({', '.join(sorted(_KERNEL_BY_NAME))}) with .variance and .lengthscales")
    active = getattr(kernel, "active_dims", None)
    if active is not None and not isinstance(active, slice):
        raise TypeError("kernels restricted to active_dims are not supported")
    variance = float(_to_numpy(kernel.variance).reshape(-1)[0])
    lengthscales = _to_numpy(kernel.lengthscales).reshape(-1)
    return _KERNEL_BY_NAME[name](variance=variance, lengthscales=lengthscales)


def likelihood_from_gpflow(likelihood) -> Gaussian:
    if isinstance(likelihood, Gaussian):
        return likelihood
    name = type(likelihood).__name__
    if name != "Gaussian" or not hasattr(likelihood, "variance"):
        raise TypeError(f"unsupported likelihood {name!r}: the reference's models use gpflow.likelihoods.Gaussian "
                        "(cggp/cli_utils.py:153,164)")
    return Gaussian(float(_to_numpy(likelihood.variance).reshape(-1)[0]))


def inducing_from_gpflow(inducing_variable, dtype=None) -> InducingPoints:
    if isinstance(inducing_variable, InducingPoints):
        return inducing_variable
    Z = getattr(inducing_variable, "Z", inducing_variable)
    if not isinstance(Z, torch.Tensor) and not (hasattr(Z, "__dlpack__") and not type(Z).__module__.startswith("numpy")):
        Z = _to_numpy(Z) if hasattr(Z, "numpy") and not isinstance(Z, np.ndarray) else np.asarray(Z)
    return InducingPoints(_lib.as_device_tensor(Z, dtype))


def from_gpflow(obj, dtype=None):
    """Convert a GPflow kernel, likelihood or inducing variable (or a tuple / list of them) to the cggp_b200 object the
    models take.  Objects of this package pass through unchanged."""
    if isinstance(obj, (tuple, list)):
        return type(obj)(from_gpflow(o, dtype) for o in obj)
    if isinstance(obj, (Stationary, Gaussian, InducingPoints)):
        return obj
    name = type(obj).__name__
    if name in _KERNEL_BY_NAME:
        return kernel_from_gpflow(obj)
    if name == "Gaussian":
        return likelihood_from_gpflow(obj)
    if hasattr(obj, "Z"):
        return inducing_from_gpflow(obj, dtype)
    raise TypeError(f"from_gpflow: cannot convert {name!r}")
