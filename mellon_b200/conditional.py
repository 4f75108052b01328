"""Posterior predictors of the conditioned Gaussian process (``mellon/conditional.py``).

Three families, as in the reference:

* ``FullConditional``               — no landmarks; ``weights = L^-T L^-1 (y - mu)``       (conditional.py:233-264)
* ``LandmarksConditional``          — rank < #landmarks; the ``A A^T`` Gram solve           (conditional.py:513-547, 57-66)
* ``LandmarksConditionalCholesky``  — rank == #landmarks; ``weights = L^-T z``              (conditional.py:805-818)

each as a plain ``Predictor`` and as a ``PredictorTime`` (``...Time``).  The O(N M) / O(N M^2)
arithmetic of ``__init__`` and every evaluation (``_mean``, ``_covariance``,
``_mean_covariance``) runs on the GPU through :mod:`mellon_b200.backend`: ``_mean`` is the fused
covariance + mat-vec kernel K7, which never materialises the (queries x landmarks) matrix.

The regression side of the same predictors (FunctionEstimator) runs on the same
kernels: multi-output ``y``, per-feature ``sigma`` of shape ``(p,)`` / ``(1, p)`` (one M x M or
N x N factorisation per output, as the reference's ``vmap``), leverage (``_leverage``), HC3 corrected
residuals and the smoothed observation variance (``_compute_obs_variance`` / ``_obs_variance``).
Still refused with ``NotImplementedError`` rather than silently computing something else: a full
noise covariance matrix as ``sigma``, the per-observation-per-feature ``(n, p)`` form, and
``obs_variance`` on the density path's ``LandmarksConditionalCholesky``.
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
        f"{what} is not part of the path mellon_b200 accelerates; use a scalar or per-feature `sigma`."
    )


def _check_obs_variance(obj):
    if not hasattr(obj, "variance_weights"):
        raise ValueError(
            "The predictor was computed without obs_variance. "
            "Recompute setting `obs_variance=True`."
        )


def _is_per_feature_sigma(sigma, y):
    """One noise level per output column of a multi-output ``y`` (conditional.py:13-36)."""
    if sigma is None or np.ndim(sigma) == 0:
        return False
    shape, yshape = np.shape(sigma), np.shape(y)
    if len(shape) == 2 and len(yshape) == 2 and shape[0] == 1 and shape[1] == yshape[1]:
        return True
    if len(shape) == 2 and len(yshape) == 2 and shape == yshape:
        return True
    if len(shape) == 1 and len(yshape) == 2 and shape[0] == yshape[1]:
        if shape[0] == yshape[0]:
            logger.warning(
                f"sigma length {shape[0]} matches both n_obs and n_features. "
                "Interpreting as per-feature. Pass sigma with shape (n, 1) for per-observation."
            )
        return True
    return False


def _sigma_columns(sigma):
    """Per-feature sigma as one noise level per output: ``(1, p)`` is squeezed to ``(p,)`` (a float each);
    the ``(n, p)`` form gives one n-vector per output (conditional.py:39-43 and the vmap axes of :249-252)."""
    s = np.asarray(sigma, dtype=float)
    if s.ndim == 2 and s.shape[0] == 1:
        s = s[0]
    if s.ndim == 1:
        return [float(v) for v in s]
    return [np.ascontiguousarray(s[:, g]) for g in range(s.shape[1])]


def _inverse_variance(s):
    """``1 / sigma^2`` of ``_process_sigma`` (conditional.py:155-159), a float or one value per observation.
    The reference divides by zero silently and returns NaN weights for ``sigma = 0``; here that is an error."""
    if not np.all(np.asarray(s) > 0):
        message = "The sparse conditional divides by sigma^2: `sigma` must be positive unless `y_is_mean=True`."
        logger.error(message)
        raise ValueError(message)
    return 1.0 / (s * s)


def _no_per_observation_feature(sigma, what):
    """The reference's leverage / observation-variance code cannot broadcast the (n, p) sigma form
    (conditional.py:319-330, 343-352, 391-398: its vmap slices rows where columns are meant) and raises from
    inside jax; refuse it by name instead."""
    if np.ndim(sigma) == 2 and np.shape(sigma)[0] != 1:
        message = f"{what} is not defined for a per-observation-per-feature `sigma` of shape {np.shape(sigma)}."
        logger.error(message)
        raise ValueError(message)


def _rows_sumsq(A):
    """Squared row norms of a device matrix as a host vector over ALL rows (gathered when sharded)."""
    be = get_backend()
    local = be.row_sumsq(A)
    return be.gather_rows(local, A.shape[0]) if A.sharded else local


def _shifted_chol(K, shift, jitter):
    """``cholesky(stabilize(K + shift * eye(n), jitter))`` on a copy of the device matrix K; ``shift`` is a
    float or one value per observation (conditional.py:245, 320, 347, 396)."""
    be = get_backend()
    if np.ndim(shift) == 0:
        A = be.add_diag(be.copy(K), float(shift) + jitter)
    else:
        A = be.add_diag_vec(be.copy(K), np.asarray(shift, dtype=float) + jitter)
    if be.potrf(A) > 0:
        _not_pd(jitter)
    return A


def _chol_solve(L, b):
    """``L^-T L^-1 b`` for a host vector / matrix b."""
    be = get_backend()
    return be.tri_solve(L, be.tri_solve(L, b), trans=True)


def _hc3(y, prediction, h):
    """Corrected squared residuals ``r^2 / (1 - h)^2`` (conditional.py:332-336, base_predictor.py:319-325)."""
    residual = np.asarray(y, dtype=float) - prediction
    if residual.ndim > h.ndim:
        h = h[..., None]
    return residual ** 2 / (1 - h) ** 2


def _full_leverage(K, sigma, jitter):
    """``h = 1 - sigma^2 diag((K + sigma^2 I)^-1)`` with ``diag(.) = colsum((L^-1)^2)``, per output when
    sigma is per-feature (conditional.py:312-330, 375-400).  K is the device matrix cov(x, x)."""
    be = get_backend()
    n = K.local_shape[0]
    _no_per_observation_feature(sigma, "The leverage")

    def one(s):
        Linv = be.tri_solve_dev(_shifted_chol(K, s * s, jitter), be.eye(n))
        return 1 - s * s * be.row_sumsq(be.transpose(Linv))

    if np.ndim(sigma) >= 1:
        return np.stack([one(s) for s in _sigma_columns(sigma)], axis=1)
    return one(float(sigma))


def _landmark_leverage(B, Lp, K_uu, sigma, jitter):
    """``diag(B M^-1 B^T)`` with ``M = sigma^2 K_uu + B^T B + jitter I`` (conditional.py:596-616, 660-685);
    ``K_uu = Lp Lp^T`` when the factor is at hand, else the matrix passed.  The reference forms
    ``B @ inv(M)``; M is symmetric positive definite, so here ``M = C C^T`` is factored and the leverage is
    the squared row norm of ``B C^-T`` (K4 Gram, K2, K3).  B is consumed when a single sigma is given."""
    be = get_backend()
    _no_per_observation_feature(sigma, "The leverage")
    BtB = be.gram(B)

    def one(s, Bw):
        if Lp is not None:
            M = be.gemm(Lp, Lp, trans_b=True, alpha=s * s, beta=1.0, out=be.copy(BtB))
        else:
            M = be.gemm(K_uu, be.eye(K_uu.local_shape[0]), alpha=s * s, beta=1.0, out=be.copy(BtB))
        be.add_diag(M, jitter)
        if be.potrf(M) > 0:
            _not_pd(jitter)
        return _rows_sumsq(be.trsm_right_lt(M, Bw))

    if np.ndim(sigma) >= 1:
        return np.stack([one(s, be.copy(B)) for s in _sigma_columns(sigma)], axis=1)
    return one(float(sigma), B)


def _scalar_sigma(sigma, what="sigma"):
    """``None`` -> None; scalars / 0-d arrays -> float; anything else is out of scope."""
    if sigma is None:
        return None
    if np.ndim(sigma) == 0:
        return float(sigma)
    _unsupported(f"A non-scalar `{what}` (shape {np.shape(sigma)})")


def _noise_diag(sigma, y_cov_factor, jitter):
    """Diagonal shift of ``add_variance(K, y_cov_factor)`` (util.py:296-331) for the noise forms
    ``_sigma_to_y_cov_factor`` turns into a diagonal factor (conditional.py:100-133): a scalar sigma gives
    ``eye(n) * sigma``, a vector ``diag(sigma)``; their ``M M^T`` is ``diag(sigma^2)``, topped up to ``jitter``.
    Returns a float or one value per observation."""
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
    if np.ndim(sigma) == 1:
        return np.maximum(np.asarray(sigma, dtype=float) ** 2, jitter)
    s2 = _scalar_sigma(sigma) ** 2
    return max(s2, jitter)


def _get_L(x, cov_func, jitter=DEFAULT_JITTER, diag_add=None):
    """``chol(cov(x, x) + jitter I)`` with the reference's failure contract (conditional.py:69-81);
    ``diag_add`` replaces the jitter by a noise level (a float, or one value per observation)."""
    be = get_backend()
    if diag_add is not None and np.ndim(diag_add) == 1:
        if len(diag_add) != np.shape(x)[0]:
            raise ValueError(f"`sigma` has {len(diag_add)} entries for {np.shape(x)[0]} observations.")
        L = be.add_diag_vec(be.cov(cov_func, x, x), diag_add)
        info = be.potrf(L)
    else:
        L, info = be.cov_chol(cov_func, x, jitter if diag_add is None else diag_add)
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
        (the estimator passes its ``Lp``).  ``y`` may have one column per output; with a per-feature
        ``sigma`` every output gets its own factor of ``K + sigma_g^2 I``."""
        be = get_backend()
        x = ensure_2d(x)
        original_sigma = sigma
        per_feature = _is_per_feature_sigma(sigma, y)
        r = np.asarray(y, dtype=float) - mu
        if per_feature:
            K = be.cov(cov_func, x, x)
            weights = np.stack([_chol_solve(_shifted_chol(K, s * s, jitter), r[:, g])
                                for g, s in enumerate(_sigma_columns(sigma))], axis=1)
        else:
            if L is None:
                logger.info("Recomputing covariance decomposition for predictive function.")
                if y_is_mean:
                    logger.debug("Assuming y is the mean of the GP.")
                    L = _get_L(x, cov_func, jitter)
                else:
                    logger.debug("Assuming y is not the mean of the GP.")
                    L = _get_L(x, cov_func, jitter, _noise_diag(sigma, y_cov_factor, jitter))
                    y_cov_factor, sigma = ("diagonal", sigma), None
            L = _dev(L)
            weights = _chol_solve(L, r)

        self.cov_func = cov_func
        self.x = _host(x)
        self.weights = weights
        self.mu = mu
        self.jitter = jitter
        self.sigma = original_sigma
        self.per_feature_sigma = per_feature
        self.n_input_features = x.shape[1]
        self.n_obs = x.shape[0]
        self._state_variables = {"x", "weights", "mu", "jitter", "sigma", "per_feature_sigma"}

        if obs_variance:
            self._compute_obs_variance(y, original_sigma)

        if not with_uncertainty:
            return
        if per_feature:
            # noise-free covariance: one factor of K + jitter I instead of one per output (conditional.py:287-290)
            L = _get_L(x, cov_func, jitter)
        self.L = L
        self._state_variables.add("L")
        if per_feature:
            return
        # W = L^-T L^-1 y_cov_factor (conditional.py:296-300)
        if isinstance(y_cov_factor, tuple):      # eye(n) * sigma, or diag(sigma)
            Linv = be.tri_solve_dev(L, be.eye(x.shape[0]))
            W = be.tri_solve_dev(L, Linv, trans=True)
            if np.ndim(y_cov_factor[1]) == 0:
                W = be.scale(W, float(y_cov_factor[1]))
            else:
                W = be.scale_cols(W, np.asarray(y_cov_factor[1], dtype=float))
        else:
            if y_cov_factor is None:
                _noise_diag(sigma, y_cov_factor, jitter)  # raises the reference's error
            F = be.copy(_dev(y_cov_factor))
            W = be.tri_solve_dev(L, be.tri_solve_dev(L, F), trans=True)
        self.W = W
        self._state_variables.add("W")

    def _compute_obs_variance(self, y, sigma):
        """HC3-corrected squared residuals at the training points, smoothed by a second GP with the same
        kernel and noise (``conditional.py:308-364``).  One factor of ``K + sigma^2 I + jitter I`` per
        noise level serves both the leverage and the second solve (the reference recomputes it)."""
        be = get_backend()
        if sigma is None:
            raise TypeError("`obs_variance=True` needs a numeric `sigma`.")
        _no_per_observation_feature(sigma, "The observation variance")
        K = be.cov(self.cov_func, self.x, self.x)
        n = self.x.shape[0]
        prediction = self._mean(self.x)
        levels = _sigma_columns(sigma) if np.ndim(sigma) >= 1 else [float(sigma)]
        factors = [_shifted_chol(K, s * s, self.jitter) for s in levels]
        hs = [1 - s * s * be.row_sumsq(be.transpose(be.tri_solve_dev(F, be.eye(n)))) for s, F in zip(levels, factors)]
        h = np.stack(hs, axis=1) if np.ndim(sigma) >= 1 else hs[0]
        corrected_r2 = _hc3(y, prediction, h)
        variance_mu = 0.0
        if np.ndim(sigma) >= 1:
            variance_weights = np.stack([_chol_solve(F, corrected_r2[:, g] - variance_mu)
                                         for g, F in enumerate(factors)], axis=1)
        else:
            variance_weights = _chol_solve(factors[0], corrected_r2 - variance_mu)
        self.variance_weights = variance_weights
        self.variance_mu = variance_mu
        self._corrected_r2 = corrected_r2
        self._state_variables.add("variance_weights")
        self._state_variables.add("variance_mu")

    def _mean(self, Xnew):
        return get_backend().predict_mean(self.cov_func, Xnew, self.x, self.weights, self.mu)

    def _leverage(self, Xnew, sigma):
        """Diagonal of the hat matrix at the TRAINING points — the reference's full predictor does not
        use ``Xnew`` (``conditional.py:375-400``)."""
        if sigma is None:
            raise TypeError("The leverage needs the numeric `sigma` the predictor was fitted with.")
        return _full_leverage(get_backend().cov(self.cov_func, self.x, self.x), sigma, self.jitter)

    def _obs_variance(self, Xnew):
        _check_obs_variance(self)
        return get_backend().predict_mean(self.cov_func, Xnew, self.x, self.variance_weights, self.variance_mu)

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
        all-reduce over the cell shards, and the M x M solves are replicated.  ``y`` may have one
        column per output; a per-feature ``sigma`` reuses the one Gram matrix for every output."""
        be = get_backend()
        x = ensure_2d(x)
        xu = ensure_2d(xu)
        original_sigma = sigma
        per_feature = _is_per_feature_sigma(sigma, y)
        if Lp is None:
            Lp = _get_L(xu, cov_func, jitter)
        Lp = _dev(Lp)
        At = be.lowrank_standard(cov_func, x, xu, Lp)
        r = np.asarray(y, dtype=float) - mu
        if r.ndim > 2:
            raise ValueError("Unsupported sigma configuration.")
        G = be.gram(At)
        self._At, self._G, self._Lp = At, G, Lp           # scratch of __init__, dropped below
        try:
            if per_feature:
                cols = _sigma_columns(sigma)
                weights = np.stack([self._sparse_solve(r[:, g], _inverse_variance(s))[0]
                                    for g, s in enumerate(cols)], axis=1)
                L_B = None
            else:
                scale = 1.0
                if not y_is_mean:
                    scale = self._noise_scale(sigma, r)
                weights, L_B = self._sparse_solve(r, scale)

            self.cov_func = cov_func
            self.landmarks = _host(xu)
            self.weights = weights
            self.mu = mu
            self.jitter = jitter
            self.sigma = original_sigma
            self.per_feature_sigma = per_feature
            self.n_input_features = xu.shape[1]
            self.n_obs = x.shape[0]
            self._state_variables = {"landmarks", "weights", "mu", "jitter", "sigma", "per_feature_sigma"}

            if obs_variance:
                self._compute_obs_variance(x, y, sigma)
            if not with_uncertainty:
                return
            self.L = Lp
            self._state_variables.add("L")
            if not per_feature:
                self.Cs = be.gemm(Lp, L_B)
                self._state_variables.add("Cs")
            if not y_is_mean:
                return
            if per_feature:
                _unsupported("`y_is_mean` together with a per-feature `sigma` and `with_uncertainty`")
            if y_cov_factor is None:
                _noise_diag(None, None, jitter)  # the reference fails here too (dot with None)
            # W = Lp^-T L_B^-T L_B^-1 (A y_cov_factor)   (conditional.py:581-586)
            Y = _dev(y_cov_factor)
            C = be.gemm(At, Y, trans_a=True, reduce=True)
            Z = be.tri_solve_dev(L_B, be.tri_solve_dev(L_B, C), trans=True)
            self.W = be.tri_solve_dev(Lp, Z, trans=True)
            self._state_variables.add("W")
        finally:
            del self._At, self._G, self._Lp

    @staticmethod
    def _noise_scale(sigma, r):
        """``_process_sigma`` (conditional.py:138-181), scalar branch: ``r_l = r / sigma^2, A_l = A / sigma^2``."""
        if sigma is not None and np.ndim(sigma) == 2 and np.shape(sigma) == (r.shape[0], r.shape[0]):
            _unsupported("A full noise covariance matrix as `sigma`")
        if sigma is not None and np.ndim(sigma) >= 1 and np.shape(sigma) == r.shape and r.ndim > 1:
            logger.error("Sigma as distinct noise per output is not implemented.")
            raise NotImplementedError("FunctionEstimator not implemented for multiple noises.")
        if sigma is not None and np.ndim(sigma) == 1 and np.shape(sigma) == r.shape:
            logger.info("Sigma interpreted as element-wise standard deviation.")
            return _inverse_variance(np.asarray(sigma, dtype=float))
        if sigma is None or np.ndim(sigma) != 0:
            raise ValueError("Unsupported sigma configuration.")
        logger.info("Sigma interpreted as element-wise standard deviation.")
        return _inverse_variance(float(sigma))

    def _sparse_solve(self, r, scale):
        """``_sparse_solve`` (conditional.py:57-66) with ``r_l = scale r`` and ``A_l = scale A``:
        returns (weights [host], L_B [device])."""
        be = get_backend()
        At, Lp = self._At, self._Lp
        if np.ndim(scale) == 1:
            # one noise level per observation: A_l A^T = sum_i a_i a_i^T / sigma_i^2, the Gram matrix of the
            # rows of A^T scaled by 1 / sigma_i; the right-hand side carries the weights on r instead
            LBB = be.gram(be.scale_rows(be.copy(At), np.sqrt(scale)))
            r, scale = r * scale, 1.0
        else:
            LBB = be.copy(self._G)
        if scale != 1.0:
            be.scale(LBB, scale)
        be.add_diag(LBB, 1.0)
        be.potrf(LBB)  # the reference does not check this factor for NaNs (conditional.py:63)
        L_B = LBB
        if r.ndim == 1:
            Ar = be.gemv_t(At, r) * scale
        else:
            Ar = be.gemm(At, r, trans_a=True, reduce=True).numpy() * scale
        c = be.tri_solve(L_B, Ar)
        return be.tri_solve(Lp, be.tri_solve(L_B, c, trans=True), trans=True), L_B

    def _compute_obs_variance(self, x, y, sigma):
        """HC3-corrected squared residuals at the training points, smoothed by a second sparse GP on the same
        landmarks and noise (``conditional.py:589-649``); reuses ``A^T`` and its Gram matrix."""
        be = get_backend()
        if sigma is None:
            raise TypeError("`obs_variance=True` needs a numeric `sigma`.")
        _no_per_observation_feature(sigma, "The observation variance")
        prediction = self._mean(x)
        B = be.cov(self.cov_func, x, self.landmarks, sharded=True)
        h = _landmark_leverage(B, self._Lp, None, sigma, self.jitter)
        corrected_r2 = _hc3(y, prediction, h)
        variance_mu = 0.0
        r_var = corrected_r2 - variance_mu
        if np.ndim(sigma) >= 1:
            cols = _sigma_columns(sigma)
            variance_weights = np.stack([self._sparse_solve(r_var[:, g], _inverse_variance(s))[0]
                                         for g, s in enumerate(cols)], axis=1)
        else:
            variance_weights, _ = self._sparse_solve(r_var, self._noise_scale(sigma, r_var))
        self.variance_weights = variance_weights
        self.variance_mu = variance_mu
        self._corrected_r2 = corrected_r2
        self._state_variables.add("variance_weights")
        self._state_variables.add("variance_mu")

    def _mean(self, Xnew):
        return get_backend().predict_mean(self.cov_func, Xnew, self.landmarks, self.weights, self.mu)

    def _leverage(self, Xnew, sigma):
        """``diag(B M^-1 B^T)`` with ``B = cov(Xnew, landmarks)`` — the reference builds ``B^T B`` from the
        points it is asked about, not from the training set (``conditional.py:660-685``)."""
        be = get_backend()
        if sigma is None:
            raise TypeError("The leverage needs the numeric `sigma` the predictor was fitted with.")
        # rows split across the ranks like the cells: the Gram contraction B^T B all-reduces over them
        B = be.cov(self.cov_func, Xnew, self.landmarks, sharded=True)
        if getattr(self, "L", None) is not None:
            return _landmark_leverage(B, _dev(self.L), None, sigma, self.jitter)
        return _landmark_leverage(B, None, be.cov(self.cov_func, self.landmarks, self.landmarks), sigma, self.jitter)

    def _obs_variance(self, Xnew):
        _check_obs_variance(self)
        return get_backend().predict_mean(self.cov_func, Xnew, self.landmarks, self.variance_weights,
                                          self.variance_mu)

    def _covariance(self, Xnew, diag=False):
        _check_covariance(self)
        Cs = None if self.per_feature_sigma else _dev(self.Cs)   # per-feature: noise-free (conditional.py:703-709)
        return _schur_covariance(self.cov_func, Xnew, self.landmarks, _dev(self.L), diag, Cs=Cs)

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

    def _leverage(self, Xnew, sigma):
        _unsupported("The leverage of a predictor built from a latent `pre_transformation`")

    def _obs_variance(self, Xnew):
        _check_obs_variance(self)

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
