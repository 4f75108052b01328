"""MAP inference: transform, objective, L-BFGS-B driver, Laplace, predictor factory
(``mellon/inference.py``).

The objective of ``compute_loss_func`` is an object, not a closure over JAX arrays: calling it
returns the loss (like the reference's function) and ``value_and_grad`` returns what
``jax.value_and_grad`` would — both from ONE fused pass over the device-resident ``L`` (K5).
The optimiser is the same SciPy L-BFGS-B the reference reaches through
``jaxopt.ScipyMinimize``, running on the host and asking the GPU for (loss, grad).
"""

from __future__ import annotations

import logging
from collections import namedtuple

import numpy as np
from scipy.optimize import minimize
from scipy.special import gammaln

from .backend import DeviceArray, get_backend
from .conditional import (FullConditional, FullConditionalTime, LandmarksConditional, LandmarksConditionalCholesky,
                          LandmarksConditionalCholeskyTime, LandmarksConditionalTime)
from .util import DEFAULT_JITTER, ensure_2d

logger = logging.getLogger("mellon")

DEFAULT_N_ITER = 100
DEFAULT_INIT_LEARN_RATE = 1e-1
DEFAULT_NUM_SAMPLES = 40
DEFAULT_OPTIMIZER = "L-BFGS-B"
DEFAULT_JIT = False
# what jaxopt.ScipyMinimize(method="L-BFGS-B") hands to scipy.optimize.minimize (maxiter=500, tol=None)
LBFGSB_OPTIONS = {"maxiter": 500}


class Transform:
    """``z -> L z + mu`` (inference.py:51-69, 125-139), evaluated on the device."""

    def __init__(self, mu, L):
        be = get_backend()
        # the estimators pass a scalar mean; the reference also accepts one value per cell (mu + L z broadcasts)
        self.mu_vector = None
        if np.ndim(mu) > 0 and np.size(mu) > 1:
            self.mu_vector = np.asarray(mu, dtype=float).reshape(-1)
            self.mu = 0.0
        else:
            self.mu = float(np.asarray(mu, dtype=float).reshape(-1)[0]) if np.ndim(mu) else float(mu)
        self.L = L if isinstance(L, DeviceArray) else be.upload(np.asarray(L, dtype=float), sharded=True)
        if self.mu_vector is not None and self.mu_vector.shape[0] != self.L.shape[0]:
            raise ValueError(f"mu has {self.mu_vector.shape[0]} entries, L has {self.L.shape[0]} rows.")

    def __call__(self, z):
        f = get_backend().transform(self.L, np.asarray(z, dtype=float), self.mu)
        return f if self.mu_vector is None else f + self.mu_vector


def compute_transform(mu, L):
    """Build the map z ~ N(0, I) -> f ~ N(mu, L L^T) (inference.py:125-139)."""
    return Transform(mu, L)


def _nn_constants(nn_distances, d):
    """V and Vdr of ``_nearest_neighbors`` (inference.py:83-85); d scalar or per-cell."""
    r = np.asarray(nn_distances, dtype=float)
    d = np.asarray(d, dtype=float) if np.ndim(d) else float(d)
    const = (d * np.log(np.pi) / 2) - gammaln(d / 2 + 1)
    log_r = np.log(r)  # one pass over the N distances (the reference evaluates log r twice)
    V = log_r * d + const
    Vdr = np.log(d) + ((d - 1) * log_r) + const
    return V, Vdr


class LossFunction:
    """``-(prior(z) + likelihood(transform(z)))`` (inference.py:35-48, 72-92, 167-192).

    ``loss(z)`` -> float, ``loss.value_and_grad(z)`` -> (float, ndarray),
    ``loss.hessian_diagonal(z)`` -> ndarray (closed form of the Laplace Hessian diagonal)."""

    def __init__(self, nn_distances, d, transform, k):
        if not isinstance(transform, Transform):
            raise TypeError(
                "compute_loss_func needs the transform built by compute_transform (a device-backed "
                "mellon_b200.inference.Transform)."
            )
        V, Vdr = _nn_constants(nn_distances, d)
        n = transform.L.shape[0]
        if V.shape != (n,):
            raise ValueError(f"nn_distances has shape {np.shape(nn_distances)}, but L has {n} rows.")
        self.k = int(k)
        self.transform = transform
        sum_vdr = float(np.sum(Vdr))
        if transform.mu_vector is not None:
            # per-cell mean: exp(L z + mu_i + V_i) and sum_i (L z + mu_i) — fold mu_i into V and into the constant
            V = V + transform.mu_vector
            sum_vdr += float(np.sum(transform.mu_vector))
        self._state = get_backend().objective(transform.L, V, sum_vdr, transform.mu, self.k)
        self.n_evaluations = 0

    def value_and_grad(self, z):
        self.n_evaluations += 1
        return get_backend().loss_grad(self._state, np.asarray(z, dtype=float))

    def __call__(self, z):
        return self.value_and_grad(z)[0]

    def grad(self, z):
        return self.value_and_grad(z)[1]

    def hessian_diagonal(self, z):
        return get_backend().hess_diag(self._state, np.asarray(z, dtype=float))


def compute_loss_func(nn_distances, d, transform, k):
    """Bayesian loss for the density model (inference.py:167-192)."""
    return LossFunction(nn_distances, d, transform, k)


ScipyMinimizeInfo = namedtuple(
    "ScipyMinimizeInfo",
    "fun_val success status iter_num hess_inv num_fun_eval num_jac_eval num_hess_eval",
)


def _value_and_grad(loss_func):
    if hasattr(loss_func, "value_and_grad"):
        return loss_func.value_and_grad

    def numeric(z):  # plain Python callables (tests, user experiments): central differences
        z = np.asarray(z, dtype=float)
        f0 = float(loss_func(z))
        g = np.empty_like(z)
        for i in range(z.size):
            h = 1e-6 * max(1.0, abs(z[i]))
            e = np.zeros_like(z)
            e[i] = h
            g[i] = (float(loss_func(z + e)) - float(loss_func(z - e))) / (2 * h)
        return f0, g

    return numeric


def minimize_lbfgsb(loss_func, initial_value, jit=DEFAULT_JIT):
    """Minimise with SciPy L-BFGS-B (inference.py:272-288).

    ``jaxopt.ScipyMinimize(method="L-BFGS-B")`` calls ``scipy.optimize.minimize(fun, x0,
    jac=True, tol=None, method="L-BFGS-B", options={"maxiter": 500})``; the same call is made
    here so the trajectory is the reference's up to the rounding of (loss, grad)."""
    from .backend import get_backend

    fun = _value_and_grad(loss_func)
    with get_backend().range("L-BFGS-B (SciPy on the host, K5 per evaluation)"):
        res = minimize(fun, np.asarray(initial_value, dtype=np.float64), jac=True, tol=None, method="L-BFGS-B",
                       options=dict(LBFGSB_OPTIONS))
    state = ScipyMinimizeInfo(
        fun_val=np.asarray(res.fun),
        success=res.success,
        status=res.status,
        iter_num=res.nit,
        hess_inv=getattr(res, "hess_inv", None),
        num_fun_eval=res.nfev,
        num_jac_eval=getattr(res, "njev", res.nfev),
        num_hess_eval=0,
    )
    Results = namedtuple("Results", "pre_transformation opt_state loss")
    return Results(res.x, state, float(res.fun))


def minimize_adam(loss_func, initial_value, n_iter=DEFAULT_N_ITER, init_learn_rate=DEFAULT_INIT_LEARN_RATE,
                  jit=DEFAULT_JIT):
    """Adam with the reference's exponentially decaying step (inference.py:222-269), the
    gradient coming from the same device pass as L-BFGS-B."""
    fun = _value_and_grad(loss_func)
    z = np.array(initial_value, dtype=np.float64)
    m, v = np.zeros_like(z), np.zeros_like(z)
    b1, b2, eps = 0.9, 0.999, 1e-8
    losses = []
    for i in range(n_iter):
        value, g = fun(z)
        losses.append(float(value))
        m = (1 - b1) * g + b1 * m
        v = (1 - b2) * np.square(g) + b2 * v
        mhat = m / (1 - b1 ** (i + 1))
        vhat = v / (1 - b2 ** (i + 1))
        z = z - np.exp(-1e-2 * i) * init_learn_rate * mhat / (np.sqrt(vhat) + eps)
    Results = namedtuple("Results", "pre_transformation opt_state losses")
    return Results(z, (z, m, v), np.stack(losses))


DEFAULT_NUM_SAMPLES = 40


def run_advi(loss_func, initial_parameters, n_iter=DEFAULT_N_ITER, init_learn_rate=DEFAULT_INIT_LEARN_RATE,
             nsamples=DEFAULT_NUM_SAMPLES, jit=DEFAULT_JIT):
    """Mean-field Gaussian variational inference with Adam (inference.py:768-876).

    The reference differentiates a Monte-Carlo ELBO with JAX; with the reparametrisation
    ``z = mean + exp(log_std) eps`` its gradients are ``E[grad loss(z)]`` for the mean and
    ``E[grad loss(z) eps std] - 1`` for ``log_std``, so every sample is one device pass of the same fused
    loss + gradient kernel L-BFGS-B uses.  As in the reference the sampler is re-keyed with the iteration
    number, the initial ``log_std`` is 0 and the step decays as ``exp(-0.01 t)``; the random streams are
    NumPy's, not threefry's, so individual ELBO values differ from a JAX run (a stochastic optimiser)."""
    fun = _value_and_grad(loss_func)
    mean = np.array(initial_parameters, dtype=np.float64)
    log_std = np.zeros_like(mean)
    params = [mean, log_std]
    m = [np.zeros_like(mean), np.zeros_like(mean)]
    v = [np.zeros_like(mean), np.zeros_like(mean)]
    b1, b2, eps_adam = 0.9, 0.999, 1e-8
    half_log_2pi = 0.5 * np.log(2.0 * np.pi)
    objective_values = []
    for t in range(n_iter):
        eps = np.random.default_rng(t).standard_normal((nsamples,) + mean.shape)
        std = np.exp(params[1])
        g_mean = np.zeros_like(mean)
        g_log_std = np.zeros_like(mean)
        total = 0.0
        for s in range(nsamples):
            value, g = fun(params[0] + std * eps[s])
            g = np.asarray(g, dtype=np.float64)
            log_q = float(np.sum(-0.5 * eps[s] ** 2 - params[1] - half_log_2pi))
            total += float(value) + log_q          # -(logprob - log q) with logprob = -loss
            g_mean += g
            g_log_std += g * eps[s] * std
        objective_values.append(total / nsamples)
        grads = [g_mean / nsamples, g_log_std / nsamples - 1.0]
        step = np.exp(-1e-2 * t) * init_learn_rate
        for i in range(2):
            m[i] = (1 - b1) * grads[i] + b1 * m[i]
            v[i] = (1 - b2) * np.square(grads[i]) + b2 * v[i]
            mhat = m[i] / (1 - b1 ** (t + 1))
            vhat = v[i] / (1 - b2 ** (t + 1))
            params[i] = params[i] - step * mhat / (np.sqrt(vhat) + eps_adam)
    Results = namedtuple("Results", "pre_transformation pre_transformation_std losses")
    return Results(params[0], np.exp(params[1]), objective_values)


def compute_laplace_std(loss_func, pre_transformation, jit=DEFAULT_JIT):
    """Diagonal Laplace posterior std ``1 / sqrt(max(diag(Hessian), 1e-8))`` (inference.py:291-338).

    For the density objective the Hessian is ``I + L^T diag(A) L`` and its diagonal is a
    weighted column sum of squares — one O(N r) pass (K6) instead of the reference's r
    Hessian-vector products.  Other callables get central second differences."""
    z = np.asarray(pre_transformation, dtype=float)
    if hasattr(loss_func, "hessian_diagonal"):
        h_diag = np.asarray(loss_func.hessian_diagonal(z))
    else:
        f0 = float(loss_func(z))
        h_diag = np.empty_like(z)
        for i in range(z.size):
            h = 1e-4 * max(1.0, abs(z[i]))
            e = np.zeros_like(z)
            e[i] = h
            h_diag[i] = (float(loss_func(z + e)) - 2 * f0 + float(loss_func(z - e))) / h ** 2
    h_diag = np.maximum(h_diag, 1e-8)
    stds = 1.0 / np.sqrt(h_diag)
    logger.info(
        "Laplace approximation: Hessian diagonal range [%.3e, %.3e], std range [%.3e, %.3e].",
        float(np.min(h_diag)), float(np.max(h_diag)), float(np.min(stds)), float(np.max(stds)),
    )
    return stds


def compute_log_density_x(pre_transformation, transform):
    """Log density at the training points: ``transform(z)`` (inference.py:341-354)."""
    return transform(pre_transformation)


def compute_parameter_cov_factor(pre_transformation_std, L):
    """``L * std[None, :]`` (inference.py:357-372)."""
    be = get_backend()
    Ld = L if isinstance(L, DeviceArray) else be.upload(np.asarray(L, dtype=float), sharded=True)
    out = be.copy_cols(Ld, 0, Ld.local_shape[1])
    return be.scale_cols(out, np.asarray(pre_transformation_std, dtype=float))


def _pick_conditional(landmarks, pre_transformation):
    if landmarks is None:
        return "full"
    if pre_transformation is not None and np.shape(pre_transformation)[0] == landmarks.shape[0]:
        return "cholesky"
    return "landmarks"


def _build(classes, x, landmarks, pre_transformation, pre_transformation_std, y, mu, cov_func, L, Lp, sigma,
           jitter, y_is_mean, with_uncertainty, time_variant, **extra):
    Full, Landmarks, Cholesky = classes
    kind = _pick_conditional(landmarks, pre_transformation)
    want_factor = pre_transformation_std is not None and (with_uncertainty or time_variant)
    if kind == "full":
        logger.debug("Using FullConditional GP.")
        y_cov_factor = compute_parameter_cov_factor(pre_transformation_std, L) if want_factor else None
        return Full(x, y, mu, cov_func, Lp, sigma=sigma, jitter=jitter, y_cov_factor=y_cov_factor,
                    y_is_mean=y_is_mean, with_uncertainty=with_uncertainty, **extra)
    landmarks = ensure_2d(landmarks)
    if kind == "cholesky":
        logger.debug("Using LandmarksConditionalCholesky GP.")
        if pre_transformation_std is not None and sigma is not None and np.any(np.asarray(sigma) > 0):
            raise ValueError(
                "One can specify either `sigma` or `pre_transformation_std` "
                "to describe uncertainty, but not both."
            )
        if pre_transformation_std is not None:
            sigma = pre_transformation_std
        return Cholesky(landmarks, pre_transformation, mu, cov_func, x.shape[0], Lp, sigma=sigma, jitter=jitter,
                        y_is_mean=y_is_mean, with_uncertainty=with_uncertainty, **extra)
    logger.debug("Using LandmarksConditional GP.")
    y_cov_factor = compute_parameter_cov_factor(pre_transformation_std, L) if want_factor else None
    # the reference passes L positionally into the `L` slot and never Lp (inference.py:493-508,
    # 619-636): the landmark Cholesky factor is recomputed inside the predictor.
    return Landmarks(x, landmarks, y, mu, cov_func, sigma=sigma, jitter=jitter, y_cov_factor=y_cov_factor,
                     y_is_mean=y_is_mean, with_uncertainty=with_uncertainty, **extra)


def compute_conditional(x, landmarks, pre_transformation, pre_transformation_std, y, mu, cov_func, L, Lp=None,
                        sigma=0, jitter=DEFAULT_JITTER, y_is_mean=False, with_uncertainty=False,
                        obs_variance=False):
    """Posterior predictor conditioned on the function values at x (inference.py:375-508).

    Chooses FullConditional (no landmarks), LandmarksConditionalCholesky (z has one entry per
    landmark) or LandmarksConditional (Nystroem ranks)."""
    extra = {"obs_variance": True} if obs_variance else {}
    return _build((FullConditional, LandmarksConditional, LandmarksConditionalCholesky), x, landmarks,
                  pre_transformation, pre_transformation_std, y, mu, cov_func, L, Lp, sigma, jitter, y_is_mean,
                  with_uncertainty, time_variant=False, **extra)


def compute_conditional_times(x, landmarks, pre_transformation, pre_transformation_std, y, mu, cov_func, L, Lp,
                              sigma=0, jitter=DEFAULT_JITTER, y_is_mean=False, with_uncertainty=False):
    """Time-aware predictor (inference.py:511-636): same selection, ``...Time`` classes."""
    return _build((FullConditionalTime, LandmarksConditionalTime, LandmarksConditionalCholeskyTime), x, landmarks,
                  pre_transformation, pre_transformation_std, y, mu, cov_func, L, Lp, sigma, jitter, y_is_mean,
                  with_uncertainty, time_variant=True)
