"""Posterior predictors of the conditioned Gaussian process (``mellon/conditional.py``).

Three families, as in the reference:

* ``FullConditional``               — no landmarks; ``weights = L^-T L^-1 (y - mu)``       (conditional.py:233-264)
* ``LandmarksConditional``          — rank < #landmarks; the ``A A^T`` Gram solve           (conditional.py:513-547, 57-66)
* ``LandmarksConditionalCholesky``  — rank == #landmarks; ``weights = L^-T z``              (conditional.py:805-818)

each as a plain ``Predictor`` and as a ``PredictorTime`` (``...Time``).  The O(N M) / O(N M^2)
arithmetic of ``__init__`` and every evaluation (``_mean``, ``_covariance``,
``_mean_covariance``) runs on the GPU through :mod:`mellon_b200.backend`: ``_mean`` is the fused
covariance + mat-vec kernel K7, which never materialises the (queries x landmarks) matrix.

Outside this package's path (they belong to FunctionEstimator): leverage,
``obs_variance`` and per-feature / full-matrix ``sigma``; asking for them raises
``NotImplementedError`` rather than silently computing something else.
"""

from __future__ import annotations

import logging

import numpy as np

from .backend import DeviceArray, get_backend
from .base_predictor import Predictor, PredictorTime
from .decomposition import DEFAULT_SIGMA
from .util import DEFAULT_JITTER, ensure_2d

logger = logging.getLogger("mellon")


def _not_pd(jitter):
    message = (
        f"Covariance not positively definite with jitter={jitter}. "
        "Consider increasing the jitter for numerical stabilization."
    )
    logger.error(message)
    raise ValueError(message)


def _check_covariance(obj):
    if not hasattr(obj, "L"):
        raise ValueError(
            "The predictor was computed without covariance. Recompute setting `with_uncertainty=True.`"
        )


def _check_uncertainty(obj):
    if not hasattr(obj, "W"):
        raise ValueError(
            "The predictor was computed without uncertainty, e.g., using ADVI. "
            "Recompute setting `with_uncertainty=True.` and define `pre_transformation_std`"
            ", e.g., by using `optimizer='advi'`."
        )


def _unsupported(what):
    raise NotImplementedError(
        f"{what} belongs to mellon's FunctionEstimator path, which mellon_b200 does not accelerate "
        "(SURVEY.md §2 row 11); use a scalar `sigma`."
    )


def _scalar_sigma(sigma, what="sigma"):
    """``None`` -> None; scalars / 0-d arrays -> float; anything else is out of scope."""
    if sigma is None:
        return None
    if np.ndim(sigma) == 0:
        return float(sigma)
    _unsupported(f"A non-scalar `{what}` (shape {np.shape(sigma)})")


def _noise_diag(sigma, y_cov_factor, jitter):
    """Diagonal shift of ``add_variance(K, y_cov_factor)`` (util.py:296-331) for the scalar-noise
    case: ``_sigma_to_y_cov_factor`` turns a scalar sigma into ``eye(n) * sigma``
    (conditional.py:100-133), whose ``M M^T`` is ``sigma^2 I`` topped up to ``jitter``."""
    if sigma is None and y_cov_factor is None:
        message = (
            "No input uncertainty specified. Make sure to set `sigma` or `pre_transformation_std`, "
            'e.g., by using `optimizer="advi", to quantify uncertainty of the prediction.'
        )
        logger.error(message)
        raise ValueError(message)
    if y_cov_factor is not None:
        if sigma is not None and np.any(np.asarray(sigma) > 0):
            raise ValueError(
                "One can specify either `sigma` or `y_cov_factor` to describe input noise, but not both."
            )
        _unsupported("Conditioning on a noise factor `y_cov_factor` without `y_is_mean`")
    s2 = _scalar_sigma(sigma) ** 2
    return max(s2, jitter)


def _get_L(x, cov_func, jitter=DEFAULT_JITTER, diag_add=None):
    """``chol(cov(x, x) + jitter I)`` with the reference's failure contract (conditional.py:69-81)."""
    L, info = get_backend().cov_chol(cov_func, x, jitter if diag_add is None else diag_add)
    if info > 0:
        _not_pd(jitter)
    return L


def _host(a):
    return a.numpy() if isinstance(a, DeviceArray) else np.asarray(a, dtype=float)


def _dev(a):
    return a if isinstance(a, DeviceArray) else get_backend().upload(np.asarray(a, dtype=float))


def _mean_covariance(cov_func, Xnew, base, W, diag):
    """``cov_L = cov(Xnew, base) W``; rows' squared norms, or ``cov_L cov_L^T``
    (conditional.py:424-440, 719-735, 947-963)."""
    be = get_backend()
    cov_L = be.gemm(be.cov(cov_func, Xnew, base), _dev(W))
    if diag:
        return be.row_sumsq(cov_L)
    return be.gemm(cov_L, cov_L, trans_b=True).numpy()


def _schur_covariance(cov_func, Xnew, base, L, diag, Cs=None):
    """``k(x*, x*) - |L^-1 k(base, x*)|^2`` (+ ``|Cs^-1 k(base, x*)|^2``)
    (conditional.py:409-422, 694-717, 930-945).  The device holds ``k(x*, base) L^-T``,
    i.e. the transpose of the reference's ``A``."""
    be = get_backend()
    At = be.trsm_right_lt(L, be.cov(cov_func, Xnew, base))
    Ct = be.trsm_right_lt(Cs, be.cov(cov_func, Xnew, base)) if Cs is not None else None
    if diag:
        var = be.cov_diag(cov_func, Xnew) - be.row_sumsq(At)
        if Ct is not None:
            var = var + be.row_sumsq(Ct)
        return var
    cov = np.asarray(cov_func(Xnew, Xnew)) - be.gemm(At, At, trans_b=True).numpy()
    if Ct is not None:
        cov = cov + be.gemm(Ct, Ct, trans_b=True).numpy()
    return cov


class _FullConditional:
    def __init__(self, x, y, mu, cov_func, L=None, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER,
                 y_cov_factor=None, y_is_mean=False, with_uncertainty=False, obs_variance=False):
        """Conditioned GP without landmarks (``conditional.py:183-372``).

        ``weights = L^-T L^-1 (y - mu)`` with ``L = chol(cov(x, x) + noise)``; ``L`` may be passed
        (the estimator passes its ``Lp``)."""
        if obs_variance:
            _unsupported("`obs_variance`")
        be = get_backend()
        x = ensure_2d(x)
        original_sigma = sigma
        if L is None:
            logger.info("Recomputing covariance decomposition for predictive function.")
            if y_is_mean:
                logger.debug("Assuming y is the mean of the GP.")
                L = _get_L(x, cov_func, jitter)
            else:
                logger.debug("Assuming y is not the mean of the GP.")
                L = _get_L(x, cov_func, jitter, _noise_diag(sigma, y_cov_factor, jitter))
                y_cov_factor, sigma = ("scalar", _scalar_sigma(sigma)), None
        L = _dev(L)
        r = np.asarray(y, dtype=float) - mu
        weights = be.tri_solve(L, be.tri_solve(L, r), trans=True)

        self.cov_func = cov_func
        self.x = _host(x)
        self.weights = weights
        self.mu = mu
        self.jitter = jitter
        self.sigma = original_sigma
        self.per_feature_sigma = False
        self.n_input_features = x.shape[1]
        self.n_obs = x.shape[0]
        self._state_variables = {"x", "weights", "mu", "jitter", "sigma", "per_feature_sigma"}

        if not with_uncertainty:
            return
        self.L = L
        self._state_variables.add("L")
        # W = L^-T L^-1 y_cov_factor (conditional.py:296-300)
        if isinstance(y_cov_factor, tuple):      # eye(n) * sigma
            Linv = be.tri_solve_dev(L, be.eye(x.shape[0]))
            W = be.scale(be.tri_solve_dev(L, Linv, trans=True), y_cov_factor[1])
        else:
            if y_cov_factor is None:
                _noise_diag(sigma, y_cov_factor, jitter)  # raises the reference's error
            F = be.copy(_dev(y_cov_factor))
            W = be.tri_solve_dev(L, be.tri_solve_dev(L, F), trans=True)
        self.W = W
        self._state_variables.add("W")

    def _mean(self, Xnew):
        return get_backend().predict_mean(self.cov_func, Xnew, self.x, self.weights, self.mu)

    def _covariance(self, Xnew, diag=True):
        _check_covariance(self)
        return _schur_covariance(self.cov_func, Xnew, self.x, _dev(self.L), diag)

    def _mean_covariance(self, Xnew, diag=True):
        _check_uncertainty(self)
        return _mean_covariance(self.cov_func, Xnew, self.x, self.W, diag)


class FullConditional(_FullConditional, Predictor):
    pass


class FullConditionalTime(_FullConditional, PredictorTime):
    pass


class _LandmarksConditional:
    def __init__(self, x, xu, y, mu, cov_func, L=None, Lp=None, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER,
                 y_cov_factor=None, y_is_mean=False, with_uncertainty=False, obs_variance=False):
        """Conditioned low-rank GP, rank < number of landmarks (``conditional.py:455-658``).

        The reference forms ``A = Lp^-1 K(xu, x)`` (M x N), ``LBB = A_l A^T + I``, its Cholesky
        ``L_B`` and ``weights = Lp^-T L_B^-T L_B^-1 (A r_l)``.  Here the device holds ``A^T``
        (N x M, row-sharded like ``x``: K1 + K3), ``A A^T`` is the Gram contraction K4 with its
        all-reduce over the cell shards, and the M x M solves are replicated."""
        if obs_variance:
            _unsupported("`obs_variance`")
        be = get_backend()
        x = ensure_2d(x)
        xu = ensure_2d(xu)
        original_sigma = sigma
        if Lp is None:
            Lp = _get_L(xu, cov_func, jitter)
        Lp = _dev(Lp)
        At = be.lowrank_standard(cov_func, x, xu, Lp)
        r = np.asarray(y, dtype=float) - mu
        if r.ndim != 1:
            _unsupported("A multi-output `y`")
        scale = 1.0
        if not y_is_mean:
            # `_process_sigma` (conditional.py:139-181), scalar case: r_l = r / sigma^2, A_l = A / sigma^2
            s = _scalar_sigma(sigma)
            if s is None:
                raise ValueError("Unsupported sigma configuration.")
            logger.info("Sigma interpreted as element-wise standard deviation.")
            scale = 1.0 / (s * s)
        # `_sparse_solve` (conditional.py:57-66)
        LBB = be.gram(At)
        if scale != 1.0:
            be.scale(LBB, scale)
        be.add_diag(LBB, 1.0)
        be.potrf(LBB)  # the reference does not check this factor for NaNs (conditional.py:63)
        L_B = LBB
        Ar = be.gemv_t(At, r) * scale
        c = be.tri_solve(L_B, Ar)
        weights = be.tri_solve(Lp, be.tri_solve(L_B, c, trans=True), trans=True)

        self.cov_func = cov_func
        self.landmarks = _host(xu)
        self.weights = weights
        self.mu = mu
        self.jitter = jitter
        self.sigma = original_sigma
        self.per_feature_sigma = False
        self.n_input_features = xu.shape[1]
        self.n_obs = x.shape[0]
        self._state_variables = {"landmarks", "weights", "mu", "jitter", "sigma", "per_feature_sigma"}

        if not with_uncertainty:
            return
        self.L = Lp
        self._state_variables.add("L")
        self.Cs = be.gemm(Lp, L_B)
        self._state_variables.add("Cs")
        if not y_is_mean:
            return
        if y_cov_factor is None:
            _noise_diag(None, None, jitter)  # the reference fails here too (dot with None)
        # W = Lp^-T L_B^-T L_B^-1 (A y_cov_factor)   (conditional.py:581-586)
        Y = _dev(y_cov_factor)
        C = be.gemm(At, Y, trans_a=True, reduce=True)
        Z = be.tri_solve_dev(L_B, be.tri_solve_dev(L_B, C), trans=True)
        self.W = be.tri_solve_dev(Lp, Z, trans=True)
        self._state_variables.add("W")

    def _mean(self, Xnew):
        return get_backend().predict_mean(self.cov_func, Xnew, self.landmarks, self.weights, self.mu)

    def _covariance(self, Xnew, diag=False):
        _check_covariance(self)
        return _schur_covariance(self.cov_func, Xnew, self.landmarks, _dev(self.L), diag, Cs=_dev(self.Cs))

    def _mean_covariance(self, Xnew, diag=True):
        _check_uncertainty(self)
        return _mean_covariance(self.cov_func, Xnew, self.landmarks, self.W, diag)


class LandmarksConditional(_LandmarksConditional, Predictor):
    pass


class LandmarksConditionalTime(_LandmarksConditional, PredictorTime):
    pass


class _LandmarksConditionalCholesky:
    def __init__(self, xu, pre_transformation, mu, cov_func, n_obs, L=None, sigma=DEFAULT_SIGMA,
                 jitter=DEFAULT_JITTER, y_is_mean=False, with_uncertainty=False, obs_variance=False,
                 obs_x=None, obs_y=None):
        """Conditioned low-rank GP, rank == number of landmarks (``conditional.py:750-906``):
        ``weights = L^-T z`` with ``L = chol(cov(xu, xu) + jitter I)``."""
        if obs_variance:
            _unsupported("`obs_variance`")
        be = get_backend()
        xu = ensure_2d(xu)
        original_sigma = sigma
        if L is None:
            logger.info("Recomputing covariance decomposition for predictive function.")
            if y_is_mean:
                logger.debug("Assuming y is the mean of the GP.")
                L = _get_L(xu, cov_func, jitter)
            else:
                logger.debug("Assuming y is not the mean of the GP.")
                if sigma is not None and np.ndim(sigma) == 1:
                    # diag(sigma) as noise factor: K + diag(max(sigma^2, jitter))
                    _unsupported("A per-landmark `sigma` without `y_is_mean`")
                L = _get_L(xu, cov_func, jitter, _noise_diag(sigma, None, jitter))
        L = _dev(L)
        weights = be.tri_solve(L, np.asarray(pre_transformation, dtype=float), trans=True)

        self.cov_func = cov_func
        self.landmarks = _host(xu)
        self.weights = weights
        self.mu = mu
        self.jitter = jitter
        self.sigma = original_sigma
        self.per_feature_sigma = False
        self.n_input_features = xu.shape[1]
        self.n_obs = n_obs
        self._state_variables = {"landmarks", "weights", "mu", "jitter", "sigma", "per_feature_sigma"}

        if not with_uncertainty:
            return
        self.L = L
        self._state_variables.add("L")
        # W = L^-T diag(sigma)  (conditional.py:859-866): columns of L^-T scaled by the std
        if sigma is None:
            raise TypeError("`sigma` (or `pre_transformation_std`) is required with `with_uncertainty=True`.")
        m = xu.shape[0]
        stds = np.asarray(sigma, dtype=float)
        if stds.ndim == 0:
            stds = np.full(m, float(stds))
        Linv = be.tri_solve_dev(L, be.eye(m))
        self.W = be.scale_cols(be.transpose(Linv), stds)
        self._state_variables.add("W")

    def _mean(self, Xnew):
        return get_backend().predict_mean(self.cov_func, Xnew, self.landmarks, self.weights, self.mu)

    def _covariance(self, Xnew, diag=True):
        _check_covariance(self)
        return _schur_covariance(self.cov_func, Xnew, self.landmarks, _dev(self.L), diag)

    def _mean_covariance(self, Xnew, diag=True):
        _check_uncertainty(self)
        return _mean_covariance(self.cov_func, Xnew, self.landmarks, self.W, diag)


class LandmarksConditionalCholesky(_LandmarksConditionalCholesky, Predictor):
    pass


class LandmarksConditionalCholeskyTime(_LandmarksConditionalCholesky, PredictorTime):
    pass
