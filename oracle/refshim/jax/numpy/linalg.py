"""`jax.numpy.linalg` on NumPy, with jax's failure conventions (NaN factors instead of exceptions)."""

import numpy as _np


def _arr(x):
    from . import _as_array

    return _as_array(x)


def cholesky(a):
    """jax symmetrises its input and returns NaNs for a non-positive-definite matrix."""
    a = _np.asarray(a)
    a = (a + _np.swapaxes(a, -1, -2)) / 2
    try:
        return _arr(_np.linalg.cholesky(a))
    except _np.linalg.LinAlgError:
        return _arr(_np.full_like(a, _np.nan, dtype=float))


def eigh(a):
    w, v = _np.linalg.eigh(_np.asarray(a))
    return _arr(w), _arr(v)


def qr(a, mode="reduced"):
    q, r = _np.linalg.qr(_np.asarray(a), mode=mode)
    return _arr(q), _arr(r)


def inv(a):
    return _arr(_np.linalg.inv(_np.asarray(a)))


def norm(x, *args, **kwargs):
    return _arr(_np.asarray(_np.linalg.norm(_np.asarray(x), *args, **kwargs)))


def lstsq(a, b, rcond=None):
    out = _np.linalg.lstsq(_np.asarray(a), _np.asarray(b), rcond=rcond)
    return tuple(_arr(_np.asarray(o)) for o in out)


def slogdet(a):
    s, l = _np.linalg.slogdet(_np.asarray(a))
    return _arr(_np.asarray(s)), _arr(_np.asarray(l))


def svd(a, full_matrices=True, compute_uv=True):
    out = _np.linalg.svd(_np.asarray(a), full_matrices=full_matrices, compute_uv=compute_uv)
    return tuple(_arr(o) for o in out) if compute_uv else _arr(out)


def matrix_rank(M, rtol=None, tol=None):
    """jax's implementation: `sum(S > rtol)` — `rtol` is compared UNSCALED (the reference's own
    known-answer test, tests/test_util.py:59-80, depends on it)."""
    M = _np.asarray(M)
    S = _np.linalg.svd(M, compute_uv=False)
    if rtol is None:
        rtol = tol
    if rtol is None:
        rtol = S.max() * max(M.shape[-2:]) * _np.finfo(S.dtype).eps
    return _arr(_np.asarray(_np.sum(S > rtol)))
