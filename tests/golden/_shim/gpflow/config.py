import numpy as np

_float = [np.float64]


def default_float():
    return _float[0]
