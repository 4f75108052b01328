"""NumPy stand-in for the subset of `jax` that settylab/Mellon imports (see ../README.md)."""

import numpy as _np

from . import numpy  # noqa: F401
from . import random  # noqa: F401
from .numpy import _as_array

__version__ = "0.0-numpy-shim"


class _Config:
    def __init__(self):
        self.values = {}

    def update(self, key, value):
        self.values[key] = value


config = _Config()


def jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def _moveaxis_in(arg, axis, i):
    if axis is None:
        return arg
    return _as_array(_np.take(_np.asarray(arg), i, axis=axis))


def vmap(fun, in_axes=0, out_axes=0):
    """Loop implementation of jax.vmap (positional arguments, pytree outputs as tuples)."""

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.shape(a)[ax]
                break
        outs = [fun(*[_moveaxis_in(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):
            oax = out_axes if isinstance(out_axes, (tuple, list)) else (out_axes,) * len(outs[0])
            return tuple(_as_array(_np.stack([_np.asarray(o[k]) for o in outs], axis=oax[k]))
                         for k in range(len(outs[0])))
        return _as_array(_np.stack([_np.asarray(o) for o in outs], axis=out_axes))

    return mapped


_H = 1e-30


def _complex_step_grad(fun, x):
    x = _np.asarray(x, dtype=float)
    flat = x.reshape(-1)
    g = _np.empty_like(flat)
    for i in range(flat.size):
        z = flat.astype(complex)
        z[i] += 1j * _H
        res = _np.asarray(fun(_as_array(z.reshape(x.shape))))
        if not _np.iscomplexobj(res):
            # the function drops the imaginary part (e.g. it evaluates on a float64 device): no complex step
            raise TypeError("function is not complex-differentiable")
        g[i] = _np.imag(res) / _H
    return g.reshape(x.shape)


def _central_grad(fun, x):
    x = _np.asarray(x, dtype=float)
    flat = x.reshape(-1)
    g = _np.empty_like(flat)
    for i in range(flat.size):
        h = 1e-6 * max(1.0, abs(flat[i]))
        e = _np.zeros_like(flat)
        e[i] = h
        g[i] = (float(fun(_as_array((flat + e).reshape(x.shape)))) -
                float(fun(_as_array((flat - e).reshape(x.shape))))) / (2 * h)
    return g.reshape(x.shape)


def _grad_of(fun, x):
    try:
        g = _complex_step_grad(fun, x)
        if _np.all(_np.isfinite(g)):
            return g
    except (TypeError, ValueError):
        pass
    return _central_grad(fun, x)


def grad(fun, argnums=0, has_aux=False):
    def g(x, *args, **kwargs):
        return _as_array(_grad_of(lambda v: fun(v, *args, **kwargs), x))

    return g


def value_and_grad(fun, argnums=0, has_aux=False):
    def vg(x, *args, **kwargs):
        f = lambda v: fun(v, *args, **kwargs)  # noqa: E731
        val = f(_as_array(_np.asarray(x, dtype=float)))
        return _as_array(_np.real(_np.asarray(val))), _as_array(_grad_of(f, x))

    return vg


def jvp(fun, primals, tangents):
    (x,), (v,) = primals, tangents
    x, v = _np.asarray(x, dtype=float), _np.asarray(v, dtype=float)
    h = 1e-5 * max(1.0, float(_np.max(_np.abs(x))))
    out = _np.asarray(fun(_as_array(x)))
    d = (_np.asarray(fun(_as_array(x + h * v))) - _np.asarray(fun(_as_array(x - h * v)))) / (2 * h)
    return _as_array(out), _as_array(d)


def jacfwd(fun, argnums=0):
    def jac(x, *args, **kwargs):
        inner = fun
        fun_ = lambda v: inner(v, *args, **kwargs)  # noqa: E731
        return _jac(fun_, x)

    return jac


def _jac(fun, x):
    if True:
        x = _np.asarray(x, dtype=float)
        out0 = _np.asarray(fun(_as_array(x)))
        cols = []
        for i in range(x.size):
            e = _np.zeros(x.size)
            e[i] = 1.0
            e = e.reshape(x.shape)
            try:
                import warnings as _warnings

                with _warnings.catch_warnings():
                    _warnings.simplefilter("ignore")
                    res = _np.asarray(fun(_as_array(x.astype(complex) + 1j * _H * e)))
                if not _np.iscomplexobj(res):
                    raise TypeError("function is not complex-differentiable")
                col = _np.imag(res) / _H
            except (TypeError, ValueError):
                h = 1e-6
                col = (_np.asarray(fun(_as_array(x + h * e))) - _np.asarray(fun(_as_array(x - h * e)))) / (2 * h)
            cols.append(col)
        return _as_array(_np.stack(cols, axis=-1).reshape(out0.shape + x.shape))


jacrev = jacfwd
