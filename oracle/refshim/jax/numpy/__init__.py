"""`jax.numpy` on NumPy: every attribute resolves to NumPy's, results come back as an ndarray
subclass that carries jax's functional `.at[idx].set(v)` update."""

import sys as _sys

import numpy as _np

from . import linalg  # noqa: F401


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Upd:
            def set(self, val):
                out = _np.array(arr)
                out[idx] = val
                return _as_array(out)

            def add(self, val):
                out = _np.array(arr)
                out[idx] += val
                return _as_array(out)

        return _Upd()


class Array(_np.ndarray):
    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


ndarray = _np.ndarray  # isinstance(x, jnp.ndarray) is true for any array


def _as_array(x):
    if isinstance(x, _np.ndarray) and not isinstance(x, Array):
        return x.view(Array)
    if isinstance(x, tuple):
        return tuple(_as_array(v) for v in x)
    return x


def _wrap(f):
    def g(*args, **kwargs):
        return _as_array(f(*args, **kwargs))

    g.__name__ = getattr(f, "__name__", "wrapped")
    g.__doc__ = getattr(f, "__doc__", None)
    return g


def array(obj, dtype=None, **kw):
    return _as_array(_np.array(obj, dtype=dtype, **kw))


def asarray(obj, dtype=None, **kw):
    # jax returns the SAME object for an array that already has the dtype (the reference's
    # `self.x is not x` identity rule, base_model.py:200-204, relies on it)
    if isinstance(obj, Array) and (dtype is None or obj.dtype == _np.dtype(dtype)):
        return obj
    return _as_array(_np.asarray(obj, dtype=dtype, **kw))


def isscalar(x):
    return _np.isscalar(x) or (isinstance(x, _np.ndarray) and x.ndim == 0)


pi = _np.pi
inf = _np.inf
nan = _np.nan
newaxis = None
float64 = _np.float64
float32 = _np.float32
int32 = _np.int32
int64 = _np.int64
bool_ = _np.bool_


def __getattr__(name):
    obj = getattr(_np, name)
    if callable(obj) and not isinstance(obj, type):
        w = _wrap(obj)
        setattr(_sys.modules[__name__], name, w)
        return w
    return obj
