import numpy as _np
import scipy.linalg as _sl

from ..numpy import _as_array


def solve_triangular(a, b, trans=0, lower=False, unit_diagonal=False, overwrite_b=False, check_finite=True):
    a, b = _np.asarray(a), _np.asarray(b)
    if a.size == 0 or b.size == 0:
        return _as_array(_np.array(b, dtype=float))
    if not (_np.all(_np.isfinite(a)) and _np.all(_np.isfinite(b))):
        return _as_array(_np.full(b.shape, _np.nan))
    return _as_array(_sl.solve_triangular(a, b, trans=trans, lower=lower, unit_diagonal=unit_diagonal))


def solve(a, b, lower=False, assume_a="gen", **kw):
    return _as_array(_sl.solve(_np.asarray(a), _np.asarray(b), lower=lower, assume_a=assume_a))


def cholesky(a, lower=False):
    from ..numpy.linalg import cholesky as _c

    L = _c(a)
    return L if lower else _as_array(_np.swapaxes(L, -1, -2))
