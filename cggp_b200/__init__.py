"""cggp_b200: B200-native (sm_100a) implementation of the conjugate-gradient hot path of
awav/conjugate-gradient-sparse-gp (CDGP / SGPR models).  Module names mirror the reference's
(``conjugate_gradient``, ``models``, ``distance``, ``utils``, ``selection``, ``covertree``); compute goes through the C ABI
``libcggp_b200.so`` (include/cggp_b200.h).  Importing the package does not need a GPU; using it does."""
from . import _lib, selection, sharding  # noqa: F401
from .covertree import CoverTree, covertree_update_inducing_parameters  # noqa: F401
from .conjugate_gradient import (BlockPreconditioner, CGPreconditioner, ConjugateGradient,  # noqa: F401
                                 DensePreconditioner, EyePreconditioner, conjugate_gradient)
from .distance import create_distance_fn, euclid_distance  # noqa: F401
from .kernels import (Gaussian, InducingPoints, Kuf, Kuu, Matern12, Matern32, Matern52, SquaredExponential,  # noqa: F401
                      prepare_points)
from .gpflow_adapter import from_gpflow  # noqa: F401
from .models import CGGP, SGPR, ClusterGP, LpSVGP, eval_logdet  # noqa: F401
from .operators import DenseOperator, SGPROperator  # noqa: F401
from .prediction import batch_posterior_computation, test_metrics  # noqa: F401
from .utils import add_diagonal  # noqa: F401


def cdgp_class(kernel, likelihood, iv, error_threshold: float = 1e-6, **kwargs):
    """cggp/cli_utils.py:439-441."""
    return CGGP(kernel, likelihood, iv, ConjugateGradient(error_threshold), **kwargs)


def sgpr_class(train_data, kernel, likelihood, iv, **kwargs):
    """cggp/cli_utils.py:444-446."""
    from .gpflow_adapter import likelihood_from_gpflow

    return SGPR(train_data, kernel, iv, noise_variance=likelihood_from_gpflow(likelihood).variance, **kwargs)
