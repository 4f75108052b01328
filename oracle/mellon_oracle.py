"""CPU oracle: NumPy/SciPy float64 restatement of settylab/Mellon's sparse-GP density hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mellon_b200/`` imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may.  It is the checker, never the product.

Pinning status: PINNED to the reference's own outputs
-----------------------------------------------------
The reference (JAX-CPU, float64) cannot be imported as is in this image (no jax / jaxlib /
jaxopt / pynndescent, no network).  ``oracle/refshim`` provides NumPy stand-ins for exactly the
API subset the reference imports, and with them the UNMODIFIED reference source is executed
from ``/root/reference`` (``oracle/make_golden.py``):

* the reference's own golden-vector tests (``tests/test_reference_results.py``, inputs from
  ``jax.random.PRNGKey(42)`` through the threefry2x32 restatement in ``refshim/jax/random.py``)
  pass on that stack, as do 198 of the 203 tests of the reference suite that do not need ADVI /
  the dimensionality estimator (the rest exercise jax autodiff features outside this path);
* its outputs on seeded NumPy inputs are committed as ``tests/golden/reference_*.npz`` and this
  oracle reproduces every one of them (``tests/test_golden_reference.py``): kernels and
  decompositions at rounding level, loss / gradient / Laplace std at fixed inputs, end-to-end
  log densities and predictions for every gp_type at the default L-BFGS-B stop and at convergence;
* every known-answer table of the reference's tests for this path is asserted in
  ``tests/test_oracle_reference_tables.py`` (that is how ``matrix_rank``'s unscaled ``rtol`` was
  caught).

This file is a line-by-line restatement, each function citing the reference ``file:line`` it
follows (paths relative to ``/root/reference/mellon``); unlike the stand-in stack it has no
dependency on ``/root/reference`` and travels to the GPU box.

Third-party arithmetic the reference reaches through un-vendored, un-pinned deps
(``pyproject.toml:20-29``): jaxopt.ScipyMinimize -> SciPy L-BFGS-B (installed SciPy is
called with the same options), sklearn Ridge / k_means (installed sklearn is called),
pynndescent.NNDescent (absent; exact 1-NN is used, which is what the reference's
known-answer tests expect), jaxlib LAPACK -> NumPy/SciPy LAPACK.
"""

from __future__ import annotations

import math
from collections import namedtuple
from enum import Enum

import numpy as np
from scipy.linalg import solve_triangular
from scipy.optimize import minimize
from scipy.special import gammaln

DEFAULT_JITTER = 1e-6          # util.py:48
DEFAULT_RANK_TOL = 5e-1        # util.py:49
DEFAULT_RANK = 0.99            # decomposition.py:17
DEFAULT_SIGMA = 0              # decomposition.py:18
DEFAULT_N_LANDMARKS = 5000     # parameters.py:53
DEFAULT_RANDOM_SEED = 42       # parameters.py:54


# --------------------------------------------------------------------------------------
# util.py
# --------------------------------------------------------------------------------------
def ensure_2d(X):
    """util.py:135-147"""
    X = np.asarray(X)
    return np.atleast_2d(X.T).T


def select_active_dims(x, active_dims):
    """util.py:150-171 — None / int / slice / list / bool mask, applied on the last axis."""
    if active_dims is not None:
        if np.isscalar(active_dims):
            active_dims = [active_dims]
        x = x[..., active_dims]
    return x


def stabilize(A, jitter=DEFAULT_JITTER):
    """util.py:269-293 — A + eye(n) * jitter."""
    return A + np.eye(A.shape[0]) * jitter


def add_variance(K, M=None, jitter=DEFAULT_JITTER):
    """util.py:296-331"""
    if M is None:
        return stabilize(K, jitter)
    if np.isscalar(M):
        return K + np.eye(K.shape[0]) * max(jitter, M ** 2)
    noise = M.dot(M.T)
    dn = np.diag(noise)
    diff = np.where(dn < jitter, jitter - dn, 0)
    return K + noise + np.diag(diff)


def mle(nn_distances, d):
    """util.py:334-348"""
    return gammaln(d / 2 + 1) - (d / 2) * np.log(np.pi) - d * np.log(nn_distances)


def distance(x, y):
    """util.py:351-366 — expansion form, +1e-12 INSIDE the sqrt, clamp at 0."""
    xx = np.sum(x * x, axis=1)[:, None]
    yy = np.sum(y * y, axis=1)[None, :]
    xy = np.tensordot(x, y, (1, 1))
    sq = xx - 2 * xy + yy + 1e-12
    return np.sqrt(np.maximum(sq, 0))


def matrix_rank_rtol(L, rtol):
    """util.py:461 — ``jnp.linalg.matrix_rank(L, rtol=tol)``.  The jax the reference was written
    against compares the singular values with ``rtol`` ITSELF (``sum(S > rtol)``, no scaling by
    the largest singular value); the reference's own known-answer test pins exactly that:
    singular values {3, 2, 1.5, 1, 0.4} with tol=0.5 must give rank 4 (tests/test_util.py:59-80)."""
    s = np.linalg.svd(np.asarray(L), compute_uv=False)
    if s.size == 0:
        return 0
    return int(np.sum(s > rtol))


def test_rank(L, tol=DEFAULT_RANK_TOL):
    """util.py:429-483 (the numeric part)."""
    if hasattr(L, "L") and not hasattr(L, "shape"):
        L = L.L
    L = np.asarray(L)
    if L.ndim != 2:
        raise ValueError("Matrix L must be 2D.")
    return matrix_rank_rtol(L, tol)


test_rank.__test__ = False  # not a pytest test


class GaussianProcessType(str, Enum):
    """util.py:589-667"""

    FULL = "full"
    FULL_NYSTROEM = "full_nystroem"
    SPARSE_CHOLESKY = "sparse_cholesky"
    SPARSE_NYSTROEM = "sparse_nystroem"
    FIXED = "fixed"


# --------------------------------------------------------------------------------------
# cov.py / base_cov.py
# --------------------------------------------------------------------------------------
class Covariance:
    """base_cov.py:17-112 (k / __call__ / diag / algebra)."""

    def __init__(self, active_dims=None):
        self.active_dims = active_dims

    def k(self, x, y):  # pragma: no cover - abstract
        raise NotImplementedError

    def __call__(self, x, y):
        return self.k(x, y)

    def diag(self, x):
        """base_cov.py:71-93 — k evaluated pairwise on (x_i, x_i)."""
        x = np.asarray(x)
        return np.array([self.k(x[i : i + 1], x[i : i + 1])[0, 0] for i in range(x.shape[0])])

    def __add__(self, other):
        return Add(self, other)

    __radd__ = __add__

    def __mul__(self, other):
        return Mul(self, other)

    __rmul__ = __mul__

    def __pow__(self, other):
        return Pow(self, other)


class _Stationary(Covariance):
    def __init__(self, ls=1.0, active_dims=None):
        super().__init__(active_dims)
        self.ls = ls

    def _dist(self, x, y):
        x = select_active_dims(x, self.active_dims)
        y = select_active_dims(y, self.active_dims)
        return distance(x, y)


class Matern32(_Stationary):
    def k(self, x, y):
        """cov.py:62-66"""
        r = np.sqrt(3.0) * self._dist(x, y) / self.ls
        return (r + 1) * np.exp(-r)


class Matern52(_Stationary):
    def k(self, x, y):
        """cov.py:157-161"""
        r = np.sqrt(5.0) * self._dist(x, y) / self.ls
        return (r + np.square(r) / 3 + 1) * np.exp(-r)


class ExpQuad(_Stationary):
    def k(self, x, y):
        """cov.py:255-259"""
        r = self._dist(x, y) / self.ls
        return np.exp(-np.square(r) / 2)


class Exponential(_Stationary):
    def k(self, x, y):
        """cov.py:352-356 — exp(-r/2), not exp(-r)."""
        r = self._dist(x, y) / self.ls
        return np.exp(-r / 2)


class RatQuad(_Stationary):
    def __init__(self, alpha=1.0, ls=1.0, active_dims=None):
        """cov.py:428 — alpha is the FIRST positional argument."""
        super().__init__(ls, active_dims)
        self.alpha = alpha

    def k(self, x, y):
        """cov.py:453-457"""
        r = self._dist(x, y) / self.ls
        return (np.square(r) / (2 * self.alpha) + 1) ** -self.alpha


class Linear(_Stationary):
    def k(self, x, y):
        """cov.py:551-556"""
        x = select_active_dims(x, self.active_dims)
        y = select_active_dims(y, self.active_dims)
        return np.einsum("ij,kj->ik", x, y) / self.ls


class _Pair(Covariance):
    def __init__(self, left, right, active_dims=None):
        super().__init__(active_dims)
        self.left = left
        self.right = right


class Add(_Pair):
    def k(self, x, y):
        """base_cov.py:309-315"""
        x = select_active_dims(x, self.active_dims)
        y = select_active_dims(y, self.active_dims)
        if callable(self.right):
            return self.left(x, y) + self.right(x, y)
        return self.left(x, y) + self.right


class Mul(_Pair):
    def k(self, x, y):
        """base_cov.py:375-381"""
        x = select_active_dims(x, self.active_dims)
        y = select_active_dims(y, self.active_dims)
        if callable(self.right):
            return self.left(x, y) * self.right(x, y)
        return self.left(x, y) * self.right


class Pow(_Pair):
    def k(self, x, y):
        """base_cov.py:449-453"""
        x = select_active_dims(x, self.active_dims)
        y = select_active_dims(y, self.active_dims)
        return self.left(x, y) ** self.right


# --------------------------------------------------------------------------------------
# decomposition.py
# --------------------------------------------------------------------------------------
def _sigma2(sigma, jitter):
    """decomposition.py:111-112 (and 159-160, 252-253)."""
    s2 = np.square(sigma)
    return np.where(s2 < jitter, jitter, s2)


def cholesky_nan(W):
    """jnp.linalg.cholesky semantics: symmetrise, lower factor, NaN instead of raising."""
    W = (W + W.T) / 2
    try:
        return np.linalg.cholesky(W)
    except np.linalg.LinAlgError:
        return np.full_like(W, np.nan)


def _eigendecomposition(A, rank=DEFAULT_RANK):
    """decomposition.py:23-76 — returns (s, v) ascending, top-p positive eigenpairs."""
    s, v = np.linalg.eigh(A)
    p = int(np.count_nonzero(s > 0))
    summed = np.cumsum(s[: -p - 1 : -1])
    if isinstance(rank, float):
        target = summed[-1] * rank
        p = int(np.searchsorted(summed, target))
        if p == 0:
            p = 1
    else:
        p = min(rank, p)
    return s[-p:], v[:, -p:]


def _full_rank(x, cov_func, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER):
    """decomposition.py:79-123"""
    W = stabilize(cov_func(x, x), _sigma2(sigma, jitter))
    L = cholesky_nan(W)
    if np.any(np.isnan(L)):
        raise ValueError(
            f"Covariance not positively definite with jitter={jitter}. "
            "Consider increasing the jitter for numerical stabilization."
        )
    return L


def _full_decomposition_low_rank(x, cov_func, rank=DEFAULT_RANK, sigma=DEFAULT_SIGMA,
                                 jitter=DEFAULT_JITTER):
    """decomposition.py:126-171"""
    W = stabilize(cov_func(x, x), _sigma2(sigma, jitter))
    s, v = _eigendecomposition(W, rank=rank)
    return v * np.sqrt(s)


def _standard_low_rank(x, cov_func, xu, Lp=None, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER):
    """decomposition.py:174-210 — L = (Lp^-1 K_MN)^T."""
    C = cov_func(x, xu)
    if Lp is None:
        Lp = _full_rank(xu, cov_func, sigma=sigma, jitter=jitter)
    return solve_triangular(Lp, C.T, lower=True).T


def _modified_low_rank(x, cov_func, xu, rank=DEFAULT_RANK, sigma=DEFAULT_SIGMA,
                       jitter=DEFAULT_JITTER):
    """decomposition.py:213-266 — improved Nystroem (QR + two eigh)."""
    W = stabilize(cov_func(xu, xu), _sigma2(sigma, jitter))
    C = cov_func(x, xu)
    Q, R = np.linalg.qr(C, mode="reduced")
    s, v = _eigendecomposition(W, rank=xu.shape[0])
    T = R @ v
    S, V = _eigendecomposition(T / s @ T.T, rank=rank)
    return Q @ V * np.sqrt(S)


# --------------------------------------------------------------------------------------
# parameters.py
# --------------------------------------------------------------------------------------
def compute_rank(gp_type):
    """parameters.py:88-115"""
    if gp_type in (GaussianProcessType.FULL_NYSTROEM, GaussianProcessType.SPARSE_NYSTROEM):
        return DEFAULT_RANK
    return 1.0


def compute_n_landmarks(gp_type, n_samples, landmarks):
    """parameters.py:118-172"""
    if landmarks is not None:
        return landmarks.shape[0]
    if gp_type is None or gp_type == GaussianProcessType.FIXED:
        return min(n_samples, DEFAULT_N_LANDMARKS)
    if gp_type in (GaussianProcessType.FULL, GaussianProcessType.FULL_NYSTROEM):
        return n_samples
    if gp_type in (GaussianProcessType.SPARSE_CHOLESKY, GaussianProcessType.SPARSE_NYSTROEM):
        return DEFAULT_N_LANDMARKS
    return min(n_samples, DEFAULT_N_LANDMARKS)


def compute_gp_type(n_landmarks, rank, n_samples):
    """parameters.py:175-240"""
    full_rank_of = lambda limit: (
        rank is None
        or (isinstance(rank, int) and rank >= limit)
        or (isinstance(rank, float) and rank >= 1.0)
        or rank == 0
    )
    if n_landmarks == 0 or n_landmarks >= n_samples:
        return GaussianProcessType.FULL if full_rank_of(n_samples) else GaussianProcessType.FULL_NYSTROEM
    return (GaussianProcessType.SPARSE_CHOLESKY if full_rank_of(n_landmarks)
            else GaussianProcessType.SPARSE_NYSTROEM)


def compute_landmarks(x, gp_type=None, n_landmarks=DEFAULT_N_LANDMARKS,
                      random_state=DEFAULT_RANDOM_SEED):
    """parameters.py:243-291 — sklearn k_means(n_init=1, random_state)."""
    from sklearn.cluster import k_means

    if n_landmarks == 0:
        return None
    x = ensure_2d(x)
    n = x.shape[0]
    if n_landmarks >= n:
        return x if gp_type == GaussianProcessType.FIXED else None
    return k_means(x, n_landmarks, n_init=1, random_state=random_state)[0]


def compute_nn_distances(x):
    """parameters.py:352-433 — exact 1-NN (pynndescent is approximate and absent here)."""
    from sklearn.neighbors import NearestNeighbors

    x = ensure_2d(np.asarray(x, dtype=float))
    n = x.shape[0]
    if n == 0:
        raise ValueError("Input data x is empty.")
    if n <= 1:
        raise ValueError(
            "Parameter k must be smaller than the number of samples. "
            f"Got k={1:,} with {n:,} samples."
        )
    nn = NearestNeighbors(n_neighbors=2).fit(x)
    dist, _ = nn.kneighbors(x)
    return dist[:, 1]


def validate_nn_distances(nn):
    """validation.py:528-592 — replace NaN / inf / <=0 by the minimum positive distance."""
    nn = np.asarray(nn, dtype=float)
    bad = np.isnan(nn) | np.isinf(nn) | (nn <= 0)
    if np.all(bad):
        raise ValueError("All nearest neighbor distances are invalid.")
    return np.where(~bad, nn, nn[~bad].min())


def compute_d(x):
    """parameters.py:534-542"""
    return 1 if np.ndim(x) < 2 else np.shape(x)[1]


def compute_mu(nn_distances, d):
    """parameters.py:586-599 — linear-interpolation 1st percentile minus 10."""
    return float(np.quantile(mle(nn_distances, d), 0.01)) - 10


def compute_ls(nn_distances):
    """parameters.py:602-613"""
    return float(np.exp(np.log(nn_distances).mean() + 3.0))


def compute_cov_func(cov_func_curry, ls, ls_time=None):
    """parameters.py:616-645"""
    if ls_time is not None:
        return cov_func_curry(ls=ls, active_dims=slice(None, -1)) * cov_func_curry(
            ls=ls_time, active_dims=-1
        )
    return cov_func_curry(ls=ls)


def compute_Lp(x, cov_func, gp_type=None, landmarks=None, sigma=DEFAULT_SIGMA,
               jitter=DEFAULT_JITTER):
    """parameters.py:648-714"""
    x = ensure_2d(x)
    n = x.shape[0]
    if landmarks is None:
        m, landmarks = n, x
    else:
        landmarks = ensure_2d(landmarks)
        m = landmarks.shape[0]
    if gp_type is None:
        gp_type = compute_gp_type(m, 1.0, n)
    if gp_type in (GaussianProcessType.FULL_NYSTROEM, GaussianProcessType.SPARSE_NYSTROEM):
        return None
    if gp_type == GaussianProcessType.FULL:
        return _full_rank(x, cov_func, sigma=sigma, jitter=jitter)
    return _full_rank(landmarks, cov_func, sigma=sigma, jitter=jitter)


def compute_L(x, cov_func, gp_type=None, landmarks=None, Lp=None, rank=None,
              sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER):
    """parameters.py:783-874"""
    x = ensure_2d(x)
    n = x.shape[0]
    m = n if landmarks is None else landmarks.shape[0]
    if rank is None:
        rank = compute_rank(gp_type)
    if gp_type is None:
        gp_type = compute_gp_type(m, rank, n)
    if landmarks is not None:
        landmarks = ensure_2d(landmarks)
    if gp_type == GaussianProcessType.FULL:
        return _full_rank(x, cov_func, sigma=sigma, jitter=jitter) if Lp is None else Lp
    if gp_type == GaussianProcessType.FULL_NYSTROEM:
        return _full_decomposition_low_rank(x, cov_func, rank=rank, sigma=sigma, jitter=jitter)
    if gp_type in (GaussianProcessType.SPARSE_CHOLESKY, GaussianProcessType.FIXED):
        return _standard_low_rank(x, cov_func, landmarks, Lp=Lp, sigma=sigma, jitter=jitter)
    return _modified_low_rank(x, cov_func, landmarks, rank=rank, sigma=sigma, jitter=jitter)


def compute_initial_value(nn_distances, d, mu, L):
    """parameters.py:877-896 — sklearn Ridge(alpha=1, fit_intercept=False)."""
    from sklearn.linear_model import Ridge

    target = mle(nn_distances, d) - mu
    return Ridge(fit_intercept=False).fit(np.asarray(L), target).coef_


def ridge_normal_equations(L, target):
    """What Ridge's `_solve_cholesky` computes (installed sklearn `_ridge.py`): primal
    (L^T L + I) z = L^T t when r <= N, dual L^T (L L^T + I)^-1 t otherwise."""
    from scipy.linalg import solve

    n, r = L.shape
    if r <= n:
        return solve(L.T @ L + np.eye(r), L.T @ target, assume_a="pos")
    return L.T @ solve(L @ L.T + np.eye(n), target, assume_a="pos")


# --------------------------------------------------------------------------------------
# inference.py
# --------------------------------------------------------------------------------------
def nn_constants(nn_distances, d):
    """inference.py:83-85 — V, Vdr (d scalar or per-cell vector)."""
    const = (d * np.log(np.pi) / 2) - gammaln(d / 2 + 1)
    V = np.log(nn_distances) * d + const
    Vdr = np.log(d) + ((d - 1) * np.log(nn_distances)) + const
    return V, Vdr


def compute_transform(mu, L):
    """inference.py:51-69,125-139"""
    L = np.asarray(L)
    return lambda z: L.dot(z) + mu


def compute_loss_func(nn_distances, d, transform, k):
    """inference.py:35-48,72-92,167-192"""
    V, Vdr = nn_constants(nn_distances, d)

    def loss_func(z):
        prior = -(1 / 2) * np.sum(z ** 2) - (k / 2) * np.log(2 * np.pi)
        f = transform(z)
        A = np.exp(f + V)
        B = f + Vdr
        return -(prior + np.sum(B - A))

    return loss_func


def loss_and_grad(L, nn_distances, d, mu, z, k=None):
    """Value and reverse-mode gradient of inference.py:189-190 (what jax.value_and_grad
    returns): grad = z + L^T (exp(Lz+mu+V) - 1)."""
    L = np.asarray(L)
    if k is None:
        k = z.shape[0]
    V, Vdr = nn_constants(nn_distances, d)
    f = L.dot(z) + mu
    A = np.exp(f + V)
    B = f + Vdr
    loss = 0.5 * np.sum(z ** 2) + (k / 2) * np.log(2 * np.pi) - np.sum(B - A)
    grad = z + L.T.dot(A - 1.0)
    return loss, grad


ScipyMinimizeInfo = namedtuple(
    "ScipyMinimizeInfo",
    "fun_val success status iter_num hess_inv num_fun_eval num_jac_eval num_hess_eval",
)
Results = namedtuple("Results", "pre_transformation opt_state loss")


def minimize_lbfgsb(value_and_grad, initial_value, options=None):
    """inference.py:272-288 via jaxopt.ScipyMinimize(method="L-BFGS-B") defaults:
    scipy.optimize.minimize(fun, x0, jac=True, tol=None, method="L-BFGS-B",
    options={"maxiter": 500})."""
    res = minimize(value_and_grad, np.asarray(initial_value, dtype=np.float64), jac=True,
                   tol=None, method="L-BFGS-B", options=options or {"maxiter": 500})
    state = ScipyMinimizeInfo(res.fun, res.success, res.status, res.nit,
                              getattr(res, "hess_inv", None), res.nfev,
                              getattr(res, "njev", res.nfev), 0)
    return Results(res.x, state, float(res.fun))


def hessian_diag(L, nn_distances, d, mu, z):
    """inference.py:311-317 in closed form: diag(I + L^T diag(A) L)."""
    L = np.asarray(L)
    V, _ = nn_constants(nn_distances, d)
    A = np.exp(L.dot(z) + mu + V)
    return 1.0 + np.einsum("i,ij,ij->j", A, L, L)


def hessian_diag_dense(L, nn_distances, d, mu, z):
    """The reference's route (M Hessian-vector products, inference.py:311-317) on the
    dense Hessian I + L^T diag(A) L — small M only; used to prove the closed form."""
    L = np.asarray(L)
    V, _ = nn_constants(nn_distances, d)
    A = np.exp(L.dot(z) + mu + V)
    H = np.eye(L.shape[1]) + L.T @ (A[:, None] * L)
    return np.array([e @ (H @ e) for e in np.eye(L.shape[1])])


def laplace_std_from_diag(h_diag):
    """inference.py:326-329 — clip at 1e-8, 1/sqrt."""
    return 1.0 / np.sqrt(np.maximum(h_diag, 1e-8))


def compute_laplace_std_numeric(loss_func, z, h=1e-4):
    """inference.py:291-338 for an arbitrary scalar function: central second differences
    (exact for the quadratic / quartic known-answer cases of tests/test_laplace.py)."""
    z = np.asarray(z, dtype=float)
    f0 = loss_func(z)
    out = np.empty_like(z)
    for i in range(z.size):
        e = np.zeros_like(z)
        e[i] = h
        out[i] = (loss_func(z + e) - 2 * f0 + loss_func(z - e)) / h ** 2
    return laplace_std_from_diag(out)


# --------------------------------------------------------------------------------------
# conditional.py  (weights + _mean of the three predictor families, y_is_mean path)
# --------------------------------------------------------------------------------------
def _get_L(x, cov_func, jitter=DEFAULT_JITTER, y_cov_factor=None):
    """conditional.py:69-81"""
    K = add_variance(cov_func(x, x), y_cov_factor, jitter=jitter)
    L = cholesky_nan(K)
    if np.any(np.isnan(L)):
        raise ValueError(
            f"Covariance not positively definite with jitter={jitter}. "
            "Consider increasing the jitter for numerical stabilization."
        )
    return L


def full_conditional_weights(x, y, mu, cov_func, L=None, jitter=DEFAULT_JITTER):
    """conditional.py:233-264 (y_is_mean, no sigma): weights = L^-T L^-1 (y - mu)."""
    x = ensure_2d(x)
    if L is None:
        L = _get_L(x, cov_func, jitter)
    r = y - mu
    return solve_triangular(L.T, solve_triangular(L, r, lower=True))


def landmarks_cholesky_weights(xu, pre_transformation, cov_func, L=None, jitter=DEFAULT_JITTER):
    """conditional.py:805-818: weights = solve_triangular(L.T, z)."""
    xu = ensure_2d(xu)
    if L is None:
        L = _get_L(xu, cov_func, jitter)
    return solve_triangular(L.T, pre_transformation)


def sparse_solve(Lp, A, r_l, A_l):
    """conditional.py:57-66"""
    LBB = stabilize(A_l @ A.T, 1)
    L_B = np.linalg.cholesky(LBB)
    c = solve_triangular(L_B, A @ r_l, lower=True)
    weights = solve_triangular(Lp.T, solve_triangular(L_B.T, c))
    return weights, L_B


def landmarks_conditional_weights(x, xu, y, mu, cov_func, Lp=None, jitter=DEFAULT_JITTER):
    """conditional.py:513-547 (y_is_mean => r_l, A_l = r, A)."""
    x = ensure_2d(x)
    xu = ensure_2d(xu)
    Kuf = cov_func(xu, x)
    if Lp is None:
        Lp = _get_L(xu, cov_func, jitter)
    A = solve_triangular(Lp, Kuf, lower=True)
    r = y - mu
    weights, _ = sparse_solve(Lp, A, r, A)
    return weights


def conditional_mean(Xnew, base, weights, mu, cov_func):
    """conditional.py:366-373 / 651-658 / 899-906: mu + cov(Xnew, base) @ weights."""
    return mu + cov_func(ensure_2d(Xnew), base).dot(weights)


# --------------------------------------------------------------------------------------
# function_estimator.py / conditional.py: regression on observed values y
# --------------------------------------------------------------------------------------
def is_per_feature_sigma(sigma, y):
    """conditional.py:13-36"""
    if sigma is None or np.ndim(sigma) == 0:
        return False
    sigma = np.asarray(sigma)
    if sigma.ndim == 2 and sigma.shape[0] == 1 and np.ndim(y) == 2 and sigma.shape[1] == y.shape[1]:
        return True
    if sigma.ndim == 2 and np.ndim(y) == 2 and sigma.shape == y.shape:
        return True
    return sigma.ndim == 1 and np.ndim(y) == 2 and sigma.shape[0] == y.shape[1]


def _sigma_columns(sigma, p):
    """One noise level (scalar, or an n-vector for the (n, p) form) per output column:
    ``_normalize_per_feature_sigma`` + the vmap axes of conditional.py:39-43, 249-252."""
    sigma = np.asarray(sigma, dtype=float)
    if sigma.ndim == 2 and sigma.shape[0] == 1 and sigma.shape[1] == p:
        sigma = sigma[0]
    return [sigma[:, g] if sigma.ndim == 2 else sigma[g] for g in range(p)]


def _noise_chol(K, sigma_g, jitter):
    """cholesky(stabilize(K + sigma_g**2 * eye(n), jitter))  (conditional.py:245, 320, 347)"""
    n = K.shape[0]
    return np.linalg.cholesky(stabilize(K + np.asarray(sigma_g) ** 2 * np.eye(n), jitter))


def _chol_solve(L, b):
    return solve_triangular(L.T, solve_triangular(L, b, lower=True))


def _full_leverage(K, sigma_g, jitter):
    """h = 1 - sigma^2 diag((K + sigma^2 I)^-1)  (conditional.py:319-330, 391-400)"""
    Linv = solve_triangular(_noise_chol(K, sigma_g, jitter), np.eye(K.shape[0]), lower=True)
    return 1 - np.asarray(sigma_g) ** 2 * np.sum(np.square(Linv), axis=0)


FunctionFit = namedtuple("FunctionFit", "kind base weights mu cov_func sigma jitter per_feature L Cs "
                                        "variance_weights corrected_r2")


def function_full_conditional(x, y, mu, cov_func, sigma=0.0, jitter=DEFAULT_JITTER, y_is_mean=False,
                              with_uncertainty=False, obs_variance=False):
    """``_FullConditional.__init__`` + ``_compute_obs_variance`` as FunctionEstimator reaches them
    (conditional.py:233-364; L is never passed on that route, function_estimator.py:356-371)."""
    x = ensure_2d(x)
    y = np.asarray(y, dtype=float)
    K = cov_func(x, x)
    per_feature = is_per_feature_sigma(sigma, y)
    r = y - mu
    if per_feature:
        cols = _sigma_columns(sigma, y.shape[1])
        weights = np.stack([_chol_solve(_noise_chol(K, s, jitter), r[:, g]) for g, s in enumerate(cols)], axis=1)
        L = None
    else:
        # y_is_mean: chol(K + jitter I); else K + max(sigma^2, jitter) I  (add_variance with eye(n) * sigma)
        L = np.linalg.cholesky(add_variance(K, None if y_is_mean else np.eye(x.shape[0]) * sigma, jitter))
        weights = _chol_solve(L, r)
    vw = cr2 = None
    if obs_variance:
        prediction = mu + K @ weights
        if np.ndim(sigma) >= 1:
            cols = _sigma_columns(sigma, y.shape[1])
            h = np.stack([_full_leverage(K, s, jitter) for s in cols], axis=1)
        else:
            h = _full_leverage(K, sigma, jitter)
        residual = y - prediction
        if residual.ndim > h.ndim:
            h = h[..., None]
        cr2 = residual ** 2 / (1 - h) ** 2
        if np.ndim(sigma) >= 1:
            vw = np.stack([_chol_solve(_noise_chol(K, s, jitter), cr2[:, g]) for g, s in enumerate(cols)], axis=1)
        else:
            vw = _chol_solve(_noise_chol(K, sigma, jitter), cr2)
    Lc = None
    if with_uncertainty:
        Lc = np.linalg.cholesky(stabilize(K, jitter)) if per_feature else L
    return FunctionFit("full", x, weights, mu, cov_func, sigma, jitter, per_feature, Lc, None, vw, cr2)


def _landmark_leverage(B, K_uu, sigma_g, jitter):
    """diag(B (sigma^2 K_uu + B^T B + jitter I)^-1 B^T)  (conditional.py:602-616, 674-685)"""
    M = stabilize(np.asarray(sigma_g) ** 2 * K_uu + B.T @ B, jitter)
    return np.sum((B @ np.linalg.inv(M)) * B, axis=1)


def function_landmarks_conditional(x, xu, y, mu, cov_func, sigma=0.0, jitter=DEFAULT_JITTER, y_is_mean=False,
                                   with_uncertainty=False, obs_variance=False):
    """``_LandmarksConditional.__init__`` + ``_compute_obs_variance`` (conditional.py:513-649)."""
    x, xu = ensure_2d(x), ensure_2d(xu)
    y = np.asarray(y, dtype=float)
    Kuf = cov_func(xu, x)
    per_feature = is_per_feature_sigma(sigma, y)
    Lp = _get_L(xu, cov_func, jitter)
    A = solve_triangular(Lp, Kuf, lower=True)
    r = y - mu
    L_B = None
    if per_feature:
        cols = _sigma_columns(sigma, y.shape[1])          # floats, or one n-vector per output for the (n, p) form
        weights = np.stack([sparse_solve(Lp, A, r[:, g] / s ** 2, A / s ** 2)[0] for g, s in enumerate(cols)], axis=1)
    elif y_is_mean:
        weights, L_B = sparse_solve(Lp, A, r, A)
    else:
        s2 = np.asarray(sigma, dtype=float) ** 2          # `_process_sigma`, element-wise branch (conditional.py:155-159)
        weights, L_B = sparse_solve(Lp, A, r / s2, A / s2)
    vw = cr2 = None
    if obs_variance:
        B = Kuf.T
        K_uu = Lp @ Lp.T
        prediction = mu + B @ weights
        if np.ndim(sigma) >= 1:
            h = np.stack([_landmark_leverage(B, K_uu, s, jitter) for s in cols], axis=1)
        else:
            h = _landmark_leverage(B, K_uu, sigma, jitter)
        residual = y - prediction
        if residual.ndim > h.ndim:
            h = h[..., None]
        cr2 = residual ** 2 / (1 - h) ** 2
        if np.ndim(sigma) >= 1:
            vw = np.stack([sparse_solve(Lp, A, cr2[:, g] / s ** 2, A / s ** 2)[0] for g, s in enumerate(cols)], axis=1)
        else:
            s2 = float(sigma) ** 2
            vw, _ = sparse_solve(Lp, A, cr2 / s2, A / s2)
    Cs = Lp @ L_B if (with_uncertainty and not per_feature) else None
    return FunctionFit("landmarks", xu, weights, mu, cov_func, sigma, jitter, per_feature,
                       Lp if with_uncertainty else None, Cs, vw, cr2)


def function_fit(x, y, landmarks=None, **kw):
    """``FunctionEstimator.fit`` -> ``compute_conditional`` (function_estimator.py:318-420, inference.py:375-508):
    no landmarks -> FullConditional, otherwise LandmarksConditional (there is no pre_transformation)."""
    if landmarks is None:
        return function_full_conditional(x, y, **kw)
    return function_landmarks_conditional(x, landmarks, y, **kw)


def function_leverage(fit, Xnew):
    """``_leverage`` (conditional.py:375-400: the full predictor ignores Xnew and returns the training leverage;
    :660-685: the landmark predictor builds B from Xnew alone)."""
    sigma, cov_func = fit.sigma, fit.cov_func
    p = np.shape(fit.weights)[1] if np.ndim(fit.weights) == 2 else 1
    if fit.kind == "full":
        K = cov_func(fit.base, fit.base)
        if np.ndim(sigma) >= 1:
            return np.stack([_full_leverage(K, s, fit.jitter) for s in _sigma_columns(sigma, p)], axis=1)
        return _full_leverage(K, sigma, fit.jitter)
    B = cov_func(ensure_2d(Xnew), fit.base)
    K_uu = fit.L @ fit.L.T if fit.L is not None else cov_func(fit.base, fit.base)
    if np.ndim(sigma) >= 1:
        return np.stack([_landmark_leverage(B, K_uu, s, fit.jitter) for s in _sigma_columns(sigma, p)], axis=1)
    return _landmark_leverage(B, K_uu, sigma, fit.jitter)


def function_loo_residuals_squared(fit, x, y):
    """base_predictor.py:290-325"""
    residual = np.asarray(y, dtype=float) - conditional_mean(x, fit.base, fit.weights, fit.mu, fit.cov_func)
    h = function_leverage(fit, x)
    if residual.ndim > h.ndim:
        h = h[..., None]
    return residual ** 2 / (1 - h) ** 2


def function_obs_variance(fit, Xnew):
    """``_obs_variance`` (conditional.py:402-407, 687-692); variance_mu is 0."""
    return fit.cov_func(ensure_2d(Xnew), fit.base) @ fit.variance_weights


def function_covariance(fit, Xnew, diag=True):
    """``_covariance`` (conditional.py:409-422, 694-717)."""
    Xnew = ensure_2d(Xnew)
    Kus = fit.cov_func(fit.base, Xnew)
    A = solve_triangular(fit.L, Kus, lower=True)
    if diag:
        var = fit.cov_func.diag(Xnew) - np.sum(np.square(A), axis=0)
        if fit.Cs is not None:
            var = var + np.sum(np.square(solve_triangular(fit.Cs, Kus, lower=True)), axis=0)
        return var
    cov = fit.cov_func(Xnew, Xnew) - A.T @ A
    if fit.Cs is not None:
        C = solve_triangular(fit.Cs, Kus, lower=True)
        cov = cov + C.T @ C
    return cov


# --------------------------------------------------------------------------------------
# density_estimator.py driver (prepare_inference -> run_inference -> process_inference)
# --------------------------------------------------------------------------------------
FitResult = namedtuple(
    "FitResult",
    "log_density_x pre_transformation loss opt_state L Lp landmarks mu ls d cov_func "
    "initial_value nn_distances gp_type",
)


def fit_density(x, cov_func_curry=Matern52, n_landmarks=None, rank=None, gp_type=None,
                jitter=DEFAULT_JITTER, landmarks=None, nn_distances=None, d=None, mu=None,
                ls=None, ls_factor=1, cov_func=None, Lp=None, L=None, initial_value=None,
                random_state=DEFAULT_RANDOM_SEED, timings=None, lbfgsb_options=None,
                grad_noise=0.0):
    """density_estimator.py:404-444, 494-516, 542-581 + base_model.py:371-431 with the
    default L-BFGS-B optimiser.  `timings`, when a dict, receives per-stage seconds."""
    import time

    tick = time.perf_counter
    x = np.asarray(x, dtype=float)
    n = x.shape[0]
    if landmarks is not None:
        landmarks = np.asarray(landmarks, dtype=float)
    if n_landmarks is None:
        n_landmarks = compute_n_landmarks(gp_type, n, landmarks)
    if rank is None:
        rank = compute_rank(gp_type)
    if gp_type is None:
        gp_type = compute_gp_type(n_landmarks, rank, n)
    if nn_distances is None:
        nn_distances = validate_nn_distances(compute_nn_distances(x))
    if d is None:
        d = compute_d(x)
    if mu is None:
        mu = compute_mu(nn_distances, d)
    if ls is None:
        ls = compute_ls(nn_distances) * ls_factor
    if cov_func is None:
        cov_func = compute_cov_func(cov_func_curry, ls)
    if landmarks is None:
        landmarks = compute_landmarks(x, gp_type, n_landmarks, random_state)
    t0 = tick()
    if Lp is None:
        Lp = compute_Lp(x, cov_func, gp_type, landmarks, sigma=0, jitter=jitter)
    t1 = tick()
    if L is None:
        L = compute_L(x, cov_func, gp_type, landmarks=landmarks, Lp=Lp, rank=rank, sigma=0,
                      jitter=jitter)
    t2 = tick()
    if initial_value is None:
        initial_value = compute_initial_value(nn_distances, d, mu, L)
    t3 = tick()
    k = initial_value.shape[0]
    if grad_noise:
        # reference-vs-reference noise floor: perturb (loss, grad) at rounding level
        noise_rng = np.random.default_rng(12345)

        def vg(z):
            l, g = loss_and_grad(L, nn_distances, d, mu, z, k)
            return l * (1 + grad_noise * noise_rng.standard_normal()), \
                g * (1 + grad_noise * noise_rng.standard_normal(g.shape))
    else:
        def vg(z):
            return loss_and_grad(L, nn_distances, d, mu, z, k)
    res = minimize_lbfgsb(vg, initial_value, options=lbfgsb_options)
    t4 = tick()
    log_density_x = np.asarray(L).dot(res.pre_transformation) + mu
    t5 = tick()
    if timings is not None:
        timings.update(Lp=t1 - t0, L=t2 - t1, init=t3 - t2, lbfgsb=t4 - t3, transform=t5 - t4,
                       total=t5 - t0, nit=res.opt_state.iter_num, nfev=res.opt_state.num_fun_eval)
    return FitResult(log_density_x, res.pre_transformation, res.loss, res.opt_state, L, Lp,
                     landmarks, mu, ls, d, cov_func, initial_value, nn_distances, gp_type)


def build_predictor(fit, x, jitter=DEFAULT_JITTER):
    """density_estimator.py:370-402 -> inference.py:375-508 (sigma=None, y_is_mean=True):
    returns (base_points, weights) such that mean(Xq) = mu + cov(Xq, base) @ weights."""
    x = ensure_2d(np.asarray(x, dtype=float))
    z = fit.pre_transformation
    if fit.landmarks is None:
        w = full_conditional_weights(x, fit.log_density_x, fit.mu, fit.cov_func, L=fit.Lp,
                                     jitter=jitter)
        return x, w
    xu = ensure_2d(fit.landmarks)
    if z.shape[0] == xu.shape[0]:
        return xu, landmarks_cholesky_weights(xu, z, fit.cov_func, L=fit.Lp, jitter=jitter)
    # inference.py:493-508 passes L positionally and no Lp -> Lp recomputed from landmarks
    return xu, landmarks_conditional_weights(x, xu, fit.log_density_x, fit.mu, fit.cov_func,
                                             Lp=None, jitter=jitter)


def predict_density(fit, x_train, Xq, jitter=DEFAULT_JITTER):
    base, w = build_predictor(fit, x_train, jitter)
    return conditional_mean(Xq, base, w, fit.mu, fit.cov_func)
