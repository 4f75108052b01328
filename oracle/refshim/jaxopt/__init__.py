"""`jaxopt.ScipyMinimize` stand-in: the wrapper jaxopt 0.8 puts around scipy.optimize.minimize
(`fun` -> (value, grad) via autodiff, `jac=True`, `tol=None`, `options={"maxiter": 500}`)."""

from collections import namedtuple

import numpy as _np
import scipy.optimize as _so

import jax as _jax

ScipyMinimizeInfo = namedtuple(
    "ScipyMinimizeInfo", "fun_val success status iter_num hess_inv num_fun_eval num_jac_eval num_hess_eval")
OptStep = namedtuple("OptStep", "params state")

# oracle/make_golden.py sets this to run the reference's optimiser to convergence ("tight" vectors)
EXTRA_OPTIONS = {}


class ScipyMinimize:
    def __init__(self, fun, method=None, tol=None, options=None, maxiter=500, jit=True, has_aux=False,
                 dtype=_np.float64, callback=None, value_and_grad=False):
        self.fun, self.method, self.tol, self.maxiter = fun, method, tol, maxiter
        self.options = dict(options or {})
        self.value_and_grad = value_and_grad

    def run(self, init_params, *args, **kwargs):
        vg = self.fun if self.value_and_grad else _jax.value_and_grad(lambda x: self.fun(x, *args, **kwargs))
        scalar = _np.ndim(init_params) == 0
        x0 = _np.atleast_1d(_np.asarray(init_params, dtype=float))

        def wrapped(x):
            v, g = vg(x[0] if scalar else x)
            return float(v), _np.atleast_1d(_np.asarray(g, dtype=float))

        res = _so.minimize(wrapped, x0, jac=True, tol=self.tol, method=self.method,
                           options={**self.options, "maxiter": self.maxiter, **EXTRA_OPTIONS})
        from jax.numpy import _as_array

        params = _as_array(_np.asarray(res.x[0] if scalar else res.x))
        state = ScipyMinimizeInfo(_as_array(_np.asarray(res.fun)), res.success, res.status, res.nit,
                                  getattr(res, "hess_inv", None), res.nfev, getattr(res, "njev", res.nfev), 0)
        return OptStep(params, state)
