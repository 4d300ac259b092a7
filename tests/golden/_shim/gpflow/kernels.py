from oracle.gpflow_restated import Matern12, Matern32, Matern52, SquaredExponential, Stationary as Kernel  # noqa: F401
