"""Covariance decompositions: builders of ``L`` with ``L L^T ~ K`` (``mellon/decomposition.py``).

Every function returns a :class:`~mellon_b200.backend.DeviceArray`; the N x M work never
leaves the GPU.  Row-sharded across ranks when a communicator is attached.
"""

from __future__ import annotations

import logging

import numpy as np

from .backend import DeviceArray, get_backend
from .util import DEFAULT_JITTER

DEFAULT_RANK = 0.99
DEFAULT_SIGMA = 0

logger = logging.getLogger("mellon")


def _noise_variance(sigma, jitter):
    """``sigma2 = square(sigma); sigma2 = jitter where sigma2 < jitter`` (decomposition.py:111-112)."""
    s2 = float(np.square(sigma))
    return jitter if s2 < jitter else s2


def _not_pd(jitter):
    message = (
        f"Covariance not positively definite with jitter={jitter}. "
        "Consider increasing the jitter for numerical stabilization."
    )
    logger.error(message)
    raise ValueError(message)


def _select_rank(s, rank):
    """The rank rule of ``_eigendecomposition`` (decomposition.py:50-76) on ASCENDING eigenvalues
    ``s``: returns p, the number of trailing eigenpairs to keep.  Integer arithmetic on M
    numbers — done on the host, identically to the reference."""
    s = np.asarray(s, dtype=float)
    if np.any(s <= 0):
        logger.warning(
            "Singuarity detected in covariance matrix. "
            "This can complicated prediction. Consider raising the jitter."
        )
    p = int(np.count_nonzero(s > 0))
    summed = np.cumsum(s[: -p - 1 : -1])
    if isinstance(rank, float):
        target = summed[-1] * rank
        p = int(np.searchsorted(summed, target))
        if p == 0:
            logger.warning(f"Low variance percentage {rank:%} indicated rank=0. Bumping rank to 1.")
            p = 1
    else:
        p = min(int(rank), p)
    if (isinstance(rank, float) and rank < 1) or rank < len(summed):
        frac = summed[min(p, len(summed) - 1)] / summed[-1]
        logger.info(f"Recovering {frac:%} variance in eigendecomposition.")
    return p


def _eigendecomposition(A, rank=DEFAULT_RANK):
    """Top positive eigenpairs of the symmetric matrix ``A`` (decomposition.py:23-76).

    Returns ``(s, v)``: eigenvalues (host, ascending) and eigenvectors (device, columns)."""
    be = get_backend()
    Ad = be.upload(np.array(A, dtype=float)) if not isinstance(A, DeviceArray) else A
    s, v = be.eigh(Ad)
    p = _select_rank(s, rank)
    n = s.shape[0]
    return s[n - p:], be.copy_cols(v, n - p, p)


def _full_rank(x, cov_func, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER):
    """``L = chol(cov(x, x) + max(sigma^2, jitter) I)`` (decomposition.py:79-123).

    K1 + K2 on the device.  A non-positive pivot raises the reference's ValueError (the
    reference detects the same failure as NaNs in JAX's factor)."""
    be = get_backend()
    L, info = be.cov_chol(cov_func, x, _noise_variance(sigma, jitter))
    if info > 0:
        _not_pd(jitter)
    return L


def _full_decomposition_low_rank(x, cov_func, rank=DEFAULT_RANK, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER):
    """``L = v sqrt(s)`` from the top eigenpairs of the full covariance (decomposition.py:126-171)."""
    be = get_backend()
    W = be.cov(cov_func, x, x)
    be.add_diag(W, _noise_variance(sigma, jitter))
    s, v = _eigendecomposition(W, rank=rank)
    return be.scale_cols(v, np.sqrt(s))


def _standard_low_rank(x, cov_func, xu, Lp=None, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER):
    """``L = cov(x, xu) Lp^-T`` (decomposition.py:174-210): K1 then the right-side solve K3."""
    be = get_backend()
    if Lp is None:
        Lp = _full_rank(xu, cov_func, sigma=sigma, jitter=jitter)
    return be.lowrank_standard(cov_func, x, xu, Lp)


def _modified_low_rank(x, cov_func, xu, rank=DEFAULT_RANK, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER):
    """Improved Nystroem factor (decomposition.py:213-266).

    The reference computes ``Q, R = qr(C)``, ``s, v = eigh(W)``, ``S, V = eigh(R v / s v^T R^T)``
    and returns ``Q V sqrt(S)`` — i.e. ``U_p sqrt(S_p)`` for the top eigenpairs of the Nystroem
    matrix ``C W^-1 C^T``.  With ``A = C Lp^-T`` (the standard low-rank factor, ``Lp = chol(W)``)
    that matrix is ``A A^T``, its non-zero spectrum is the spectrum of the M x M Gram ``A^T A``
    and ``U_p sqrt(S_p) = A V_p`` for the Gram's top eigenvectors ``V_p``.  So the device path is
    K1 + K3 (A), K4 (Gram + all-reduce over the cell shards), ONE M x M eigh, and one N x M x p
    product — no tall QR, no second eigh.  Columns of L are defined up to sign in both routes.
    """
    be = get_backend()
    Lp = _full_rank(xu, cov_func, sigma=sigma, jitter=jitter)
    A = be.lowrank_standard(cov_func, x, xu, Lp)
    G = be.gram(A)
    S, V = be.eigh(G)
    p = _select_rank(S, rank)
    m = S.shape[0]
    Vp = be.copy_cols(V, m - p, p)
    return be.gemm(A, Vp)
