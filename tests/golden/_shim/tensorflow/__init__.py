"""Minimal NumPy stand-in for the TensorFlow symbols that the reference's ``cggp/conjugate_gradient.py``,
``cggp/models.py`` and ``cggp/utils.py`` call, so those files run UNMODIFIED in a container without TensorFlow.
Used only by ``tests/golden/make_golden.py`` to generate golden vectors.  Every function maps 1:1 to the NumPy op of
the same meaning; nothing here implements any cggp algorithm.
"""
import types

import numpy as np
from scipy import linalg as _sl

Tensor = np.ndarray
int32 = np.int32
int64 = np.int64
float32 = np.float32
float64 = np.float64

# make_golden.py reads this: loop variables at every evaluation of a while_loop condition
while_loop_trace = []


class Variable(np.ndarray):
    def __new__(cls, value, dtype=None, shape=None, trainable=True, **kw):
        return np.array(value, dtype=dtype).view(cls)

    def assign(self, value):
        self[...] = np.asarray(value, dtype=self.dtype)
        return self


def convert_to_tensor(value, dtype=None):
    return np.asarray(value, dtype=dtype)


def constant(value, dtype=None):
    return np.asarray(value, dtype=dtype)


def cast(x, dtype):
    return np.asarray(x).astype(dtype)


def shape(x):
    return np.array(np.shape(x), dtype=np.int64)


def reduce_sum(x, axis=None, keepdims=False):
    return np.sum(x, axis=axis, keepdims=keepdims)


def reduce_any(x):
    return np.any(x)


def reduce_mean(x, axis=None):
    return np.mean(x, axis=axis)


square = np.square
sqrt = np.sqrt
logical_and = np.logical_and
where = np.where
zeros_like = np.zeros_like
ones_like = np.ones_like


def transpose(x):
    return np.transpose(np.asarray(x))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a = np.asarray(a)
    b = np.asarray(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return a @ b


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(pred) else false_fn()


def while_loop(cond_fn, body, loop_vars, **kw):
    loop_vars = list(loop_vars)
    while True:
        while_loop_trace.append(loop_vars)
        if not bool(cond_fn(*loop_vars)):
            break
        loop_vars = list(body(*loop_vars))
    return loop_vars


# custom_gradient: forward value only; the gradient closure is kept for make_golden.py
last_custom_gradient = []


def custom_gradient(fn):
    def wrapped(*args):
        out, grad = fn(*args)
        last_custom_gradient.append(grad)
        return out

    return wrapped


def function(func=None, **kw):
    return func


def _cholesky(a):
    return np.linalg.cholesky(a)


def _cholesky_solve(L, rhs):
    return _sl.cho_solve((np.asarray(L), True), np.asarray(rhs))


def _triangular_solve(L, rhs, lower=True):
    return _sl.solve_triangular(np.asarray(L), np.asarray(rhs), lower=lower)


def _set_diag(m, d):
    out = np.array(m, copy=True)
    i = np.arange(out.shape[-1])
    out[i, i] = d
    return out


linalg = types.SimpleNamespace(
    cholesky=_cholesky,
    cholesky_solve=_cholesky_solve,
    triangular_solve=_triangular_solve,
    trace=np.trace,
    diag_part=lambda m: np.diagonal(np.asarray(m)).copy(),
    set_diag=_set_diag,
    eye=lambda n, dtype=None: np.eye(int(n), dtype=dtype),
    norm=lambda x, axis=None: np.linalg.norm(x, axis=axis),
    solve=np.linalg.solve,
    logdet=lambda a: np.linalg.slogdet(a)[1],
    adjoint=lambda a: np.swapaxes(a, -1, -2),
)
math = types.SimpleNamespace(log=np.log, reciprocal=np.reciprocal, argmin=np.argmin, argmax=np.argmax, cos=np.cos,
                             sin=np.sin, truediv=np.true_divide)  # cos / sin / truediv: cggp/rff.py
data = types.SimpleNamespace(Dataset=object, experimental=types.SimpleNamespace(AUTOTUNE=-1))
concat = lambda xs, axis=0: np.concatenate(xs, axis=axis)  # noqa: E731
