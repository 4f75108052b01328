"""Host utilities mirroring ``mellon/util.py`` (names, defaults, log text).

Only O(N) scalar work and bookkeeping lives here; anything touching an N x M array goes
through :mod:`mellon_b200.backend`.
"""

from __future__ import annotations

import logging
from enum import Enum
from typing import List

import numpy as np
from scipy.special import gammaln

logger = logging.getLogger("mellon")

DEFAULT_JITTER = 1e-6      # mellon/util.py:48
DEFAULT_RANK_TOL = 5e-1    # mellon/util.py:49


# ---- (de)serialisation helpers: same wire format as mellon/util.py:69-132 -------------------
def _none_to_str(v):
    return "None" if v is None else v


def _str_to_none(v):
    return None if isinstance(v, str) and v == "None" else v


def make_serializable(x):
    """JSON-ready form; arrays keep the reference's ``"jax.numpy"`` tag for file compatibility."""
    from .backend import DeviceArray

    if isinstance(x, DeviceArray):
        x = x.numpy()
    if isinstance(x, np.ndarray):
        return {"type": "jax.numpy", "data": x.tolist()}
    if isinstance(x, np.integer):
        return int(x)
    if isinstance(x, np.floating):
        return float(x)
    if isinstance(x, slice):
        return {"type": "slice", "data": [_none_to_str(v) for v in (x.start, x.stop, x.step)]}
    if isinstance(x, dict):
        return {"type": "dict", "data": {k: make_serializable(v) for k, v in x.items()}}
    if isinstance(x, set):
        return {"type": "set", "data": [make_serializable(v) for v in x]}
    return _none_to_str(x)


def deserialize(s):
    if isinstance(s, dict):
        kind = s["type"]
        if kind == "jax.numpy":
            return np.array(s["data"])
        if kind == "slice":
            return slice(*[_str_to_none(v) for v in s["data"]])
        if kind == "dict":
            return {k: deserialize(v) for k, v in s["data"].items()}
        if kind == "set":
            return {deserialize(v) for v in s["data"]}
        return None
    return _str_to_none(s)


# ---- shapes / dims -----------------------------------------------------------------------------
def ensure_2d(X):
    """mellon/util.py:135-147 — a 1-D array becomes a column."""
    from .backend import DeviceArray

    if isinstance(X, DeviceArray):
        return X
    X = np.asarray(X)
    return np.atleast_2d(X.T).T


def select_active_dims(x, active_dims):
    """mellon/util.py:150-171"""
    if active_dims is not None:
        if np.isscalar(active_dims):
            active_dims = [active_dims]
        x = np.asarray(x)[..., active_dims]
    return x


def expand_to_inactive(values, target_shape, active_dims):
    """mellon/util.py:174-203 — scatter gradients of active dims into a zero array."""
    if active_dims is None:
        return values
    out = np.zeros(target_shape)
    if np.isscalar(active_dims):
        active_dims = [active_dims]
    out[..., active_dims] = values
    return out


def mle(nn_distances, d):
    """mellon/util.py:334-348 — nearest-neighbour maximum-likelihood log density."""
    nn_distances = np.asarray(nn_distances, dtype=float)
    d = np.asarray(d, dtype=float) if np.ndim(d) else d
    return gammaln(d / 2 + 1) - (d / 2) * np.log(np.pi) - d * np.log(nn_distances)


def distance_grad(x, eps=1e-12):
    """``y -> (distance(x, y), d distance / d y)`` with shapes (n, m) and (n, m, d) (mellon/util.py:369-428).
    The distances are the device kernel's; the gradient ``(y - x) / (distance + eps)`` is an element-wise host
    expression over an array that is d times larger than anything on the path — a derivative utility like the
    predictors' ``gradient``, not part of the accelerated path."""
    x = ensure_2d(np.asarray(x, dtype=float))

    def grad(y):
        y = ensure_2d(np.asarray(y, dtype=float))
        dist = np.asarray(distance(x, y))
        if eps != 1e-12:                       # the kernel's constant is 1e-12 (util.py:365); re-base for another eps
            dist = np.sqrt(np.maximum(dist * dist - 1e-12 + eps, 0))
        delta = y[np.newaxis, :] - x[:, np.newaxis]
        return dist, delta / (dist[..., np.newaxis] + eps)

    return grad


def set_jax_config(enable_x64=True, platform_name="cpu"):
    """Kept for source compatibility (mellon/util.py:572-586): there is no JAX to configure.  Arithmetic is float64 on
    the GPU whatever ``platform_name`` says; single precision is refused."""
    if not enable_x64:
        raise ValueError("mellon_b200 computes in float64 only.")


def distance(x, y):
    """mellon/util.py:351-366 — sqrt(max(xx - 2xy + yy + 1e-12, 0)), evaluated by the same
    fused tile kernel as the covariances (leaf kind MB_K_DISTANCE)."""
    from .cov import _Distance

    return _Distance()(x, y)


def stabilize(A, jitter=DEFAULT_JITTER):
    """mellon/util.py:283-293 — A + jitter I (host arrays; the device path uses add_diag)."""
    A = np.asarray(A)
    return A + np.eye(A.shape[0]) * jitter


def test_rank(input, tol=DEFAULT_RANK_TOL, threshold=None):
    """mellon/util.py:429-483 — approximate rank of L = #singular values > tol (the jax the
    reference targets compares with ``rtol`` unscaled; pinned by tests/test_util.py:59-80).

    The singular values of L are the square roots of the eigenvalues of the r x r Gram matrix
    L^T L, which the device already knows how to build and all-reduce (K4); no N x r SVD."""
    from .backend import DeviceArray, get_backend

    if hasattr(input, "shape"):
        L = input
    elif hasattr(input, "L"):
        L = input.L
        if L is None:
            raise AttributeError(
                "Matrix L is not found in the estimator object. Consider running `.prepare_inference()`."
            )
    else:
        raise TypeError("Input must be either a matrix or a mellon enstimator with a transformation L.")
    if len(L.shape) != 2:
        raise ValueError("Matrix L must be 2D.")

    be = get_backend()
    Ld = L if isinstance(L, DeviceArray) else be.upload(np.asarray(L, dtype=float), sharded=True)
    if L.shape[1] <= L.shape[0]:
        G = be.gram(Ld)
    else:  # wide matrix: Gram over the short side
        G = be.gemm(Ld, Ld, trans_b=True)
    ev, _ = be.eigh(G)
    sv = np.sqrt(np.clip(ev, 0.0, None))
    approx_rank = int(np.sum(sv > tol)) if sv.size else 0
    max_rank = min(L.shape)
    rank_fraction = approx_rank / max_rank
    if threshold is not None:
        if rank_fraction > threshold:
            logger.warning(
                f"High approx. rank fraction ({rank_fraction:.1%}). Consider increasing 'n_landmarks'."
            )
        else:
            logger.info(
                f"Rank fraction ({rank_fraction:.1%}, lower is better) is within acceptable range. "
                "Current settings should provide satisfactory model performance."
            )
    else:
        print(
            f"The approx. rank fraction is {rank_fraction:.1%} "
            f"({approx_rank:,} of {max_rank:,}). Lower is better."
        )
    return approx_rank


test_rank.__test__ = False


def set_verbosity(verbose: bool):
    """mellon/util.py:539-569"""
    level = logging.INFO if verbose else logging.WARNING
    logger.setLevel(level)
    logger.info(f"Logging verbosity set to {'INFO' if verbose else 'WARNING'}.")


class GaussianProcessType(str, Enum):
    """mellon/util.py:589-667 — 'full', 'full_nystroem', 'sparse_cholesky', 'sparse_nystroem', 'fixed'."""

    FULL = "full"
    FULL_NYSTROEM = "full_nystroem"
    SPARSE_CHOLESKY = "sparse_cholesky"
    SPARSE_NYSTROEM = "sparse_nystroem"
    FIXED = "fixed"

    @staticmethod
    def from_string(s, optional: bool = False):
        if s is None:
            if optional:
                return None
            logger.error("Gaussian process type must be specified but is None.")
            raise ValueError("Gaussian process type must be specified but is None.")
        if isinstance(s, GaussianProcessType):
            return s
        wanted = s.lower().replace(" ", "_")
        for member in GaussianProcessType:
            if member.value == wanted:
                logger.info(f"Gaussian Process type: {member.value}")
                return member
        for member in GaussianProcessType:
            if wanted in member.value:
                logger.warning(
                    f"Partial match found for Gaussian Process type: {member.value}. Input was: {s}"
                )
                return member
        message = f"Unknown Gaussian Process type: {s}"
        logger.error(message)
        raise ValueError(message)


def object_str(obj: object, dim_names: List[str] = None) -> str:
    """mellon/util.py:670-711 — '<array 3 cells x 2 ranks, dtype=float64>'."""
    if hasattr(obj, "shape") and hasattr(obj, "dtype"):
        dims = obj.shape
        names = list(dim_names or [])
        parts = [f"{dim:,} {name}" for dim, name in zip(dims, names)] if names else [f"{dim:,}" for dim in dims]
        for i in range(len(parts), len(dims)):
            parts.append(f"{dims[i]} dimension {i + 1}")
        return f"<array {' x '.join(parts)}, dtype={obj.dtype}>"
    return str(obj)


def object_html(obj: object, dim_names: list = None) -> str:
    """mellon/util.py:714-761"""
    import html

    if hasattr(obj, "shape") and hasattr(obj, "dtype"):
        names = list(dim_names or [])
        names += [None] * (len(obj.shape) - len(names))
        parts = [f"{dim:,} {name}" if name else f"{dim:,}" for dim, name in zip(obj.shape, names)]
        return f"<span>&lt;array {html.escape(' x '.join(parts))}, dtype={html.escape(str(obj.dtype))}&gt;</span>"
    return f"<span>{html.escape(str(obj), quote=True)}</span>"


def make_multi_time_argument(func):
    """mellon/util.py:206-266 — adds ``multi_time`` to ``func(self, x, time=None, ...)``:
    the function is evaluated once per time point and the results stacked on axis 1."""
    import functools

    @functools.wraps(func)
    def wrapper(self, x, *args, **kwargs):
        multi_time = kwargs.pop("multi_time", None)
        time = args[0] if args else kwargs.get("time", None)
        if multi_time is not None:
            if time is not None:
                raise ValueError("Both 'time' and 'multi_time' arguments were passed. Provide only one.")
            rest = args[1:]
            kwargs.pop("time", None)
            multi_time = np.atleast_1d(np.asarray(multi_time, dtype=float))
            cols = [np.asarray(func(self, x, float(t), *rest, **kwargs)) for t in multi_time]
            return np.stack(cols, axis=1)
        return func(self, x, *args, **kwargs)

    return wrapper
