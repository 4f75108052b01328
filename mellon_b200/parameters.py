"""Parameter heuristics and the ``compute_*`` entry points (``mellon/parameters.py``).

The O(N) scalar heuristics (mu, ls, d) are NumPy on the host; ``compute_Lp`` / ``compute_L`` /
``compute_initial_value`` drive the device.  Landmark (k-means) and neighbour selection stay
with scikit-learn: they are the step *before* the accelerated path and their index selection
has to be identical on both sides of a parity check.
"""

from __future__ import annotations

import logging

import numpy as np

from .backend import DeviceArray, get_backend
from .decomposition import (DEFAULT_RANK, DEFAULT_SIGMA, _full_decomposition_low_rank, _full_rank, _modified_low_rank,
                            _standard_low_rank)
from .parameter_validation import validate_normalize_parameter, validate_params
from .util import DEFAULT_JITTER, GaussianProcessType, ensure_2d, mle
from .validation import (validate_array, validate_float_or_int, validate_float_or_iterable_numerical, validate_k,
                         validate_positive_float, validate_positive_int, validate_time_x)

DEFAULT_N_LANDMARKS = 5000
DEFAULT_RANDOM_SEED = 42

logger = logging.getLogger("mellon")

_NYSTROEM = (GaussianProcessType.FULL_NYSTROEM, GaussianProcessType.SPARSE_NYSTROEM)
_LANDMARK_CHOLESKY = (GaussianProcessType.SPARSE_CHOLESKY, GaussianProcessType.FIXED)


def compute_rank(gp_type):
    """Default rank for a gp_type: 0.99 for the Nystroem types, else 1.0 (parameters.py:88-115)."""
    return DEFAULT_RANK if gp_type in _NYSTROEM else 1.0


def compute_n_landmarks(gp_type, n_samples, landmarks):
    """Number of landmarks implied by gp_type / given landmarks (parameters.py:118-172)."""
    if landmarks is not None:
        return landmarks.shape[0]
    if gp_type is None or gp_type == GaussianProcessType.FIXED:
        return min(n_samples, DEFAULT_N_LANDMARKS)
    if gp_type in (GaussianProcessType.FULL, GaussianProcessType.FULL_NYSTROEM):
        return n_samples
    if gp_type in (GaussianProcessType.SPARSE_CHOLESKY, GaussianProcessType.SPARSE_NYSTROEM):
        if n_samples <= DEFAULT_N_LANDMARKS:
            logger.warning(
                f"Gaussian Process type {gp_type} and default "
                f"number of landmarks {DEFAULT_N_LANDMARKS:,} < "
                f"number of cells {n_samples:,}. Reduce n_landmarks below "
                f"the number of cells to use {gp_type}."
            )
        return DEFAULT_N_LANDMARKS
    n_landmarks = min(n_samples, DEFAULT_N_LANDMARKS)
    logger.warning(f"Unknown Gaussian Process type {gp_type}, using default n_landmarks={n_landmarks:,}.")
    return n_landmarks


def compute_gp_type(n_landmarks, rank, n_samples):
    """Resolve the GP type from landmark count, rank and sample count (parameters.py:175-240)."""
    rank = validate_float_or_int(rank, "rank", optional=True)
    n_landmarks = validate_positive_int(n_landmarks, "n_landmarks")
    n_samples = validate_positive_int(n_samples, "n_samples")

    def is_full_rank(limit):
        return (
            rank is None
            or (isinstance(rank, int) and rank >= limit)
            or (isinstance(rank, float) and rank >= 1.0)
            or rank == 0
        )

    if n_landmarks == 0 or n_landmarks >= n_samples:
        if is_full_rank(n_samples):
            logger.info(
                "Using non-sparse Gaussian Process since n_landmarks "
                f"({n_landmarks:,}) >= n_samples ({n_samples:,}) and rank = {rank}."
            )
            return GaussianProcessType.FULL
        logger.info(
            "Using full Gaussian Process with Nyström rank reduction since n_landmarks "
            f"({n_landmarks:,}) >= n_samples ({n_samples:,}) and rank = {rank}."
        )
        return GaussianProcessType.FULL_NYSTROEM
    if is_full_rank(n_landmarks):
        logger.info(
            "Using sparse Gaussian Process since n_landmarks "
            f"({n_landmarks:,}) < n_samples ({n_samples:,}) and rank = {rank}."
        )
        return GaussianProcessType.SPARSE_CHOLESKY
    logger.info(
        "Using sparse Gaussian Process with improved Nyström rank reduction since n_landmarks "
        f"({n_landmarks:,}) >= n_samples ({n_samples:,}) and rank = {rank}."
    )
    return GaussianProcessType.SPARSE_NYSTROEM


def compute_landmarks(x, gp_type=None, n_landmarks=DEFAULT_N_LANDMARKS, random_state=DEFAULT_RANDOM_SEED):
    """k-means centroids as landmarks (parameters.py:243-291); None when n_landmarks >= n.  scikit-learn's
    ``k_means(x, k, n_init=1, random_state=seed)`` algorithm with its distance work on the device (kmeans.py)."""
    from .kmeans import k_means

    if n_landmarks == 0:
        return None
    x = ensure_2d(np.asarray(x, dtype=float))
    n = x.shape[0]
    assert n_landmarks > 1, "n_landmarks musst be larger 1 or euqual to 0"
    if n_landmarks >= n:
        if gp_type == GaussianProcessType.FIXED:
            logger.info(
                f"Gaussian process type is {gp_type} and n_landmarks={n_landmarks:,} "
                f"requested while only {n:,} datapoints are available. "
                f"Using all {n:,} datapoints as landmarks."
            )
            return x
        return None
    logger.info(f"Computing {n_landmarks:,} landmarks with k-means clustering (random_state={random_state}).")
    return k_means(x, n_landmarks, random_state=random_state)


def compute_landmarks_rescale_time(x, ls, ls_time, times=None, n_landmarks=DEFAULT_N_LANDMARKS,
                                   random_state=DEFAULT_RANDOM_SEED):
    """k-means on the time-rescaled state (parameters.py:294-349)."""
    if n_landmarks == 0:
        return None
    ls = validate_positive_float(ls, "ls")
    ls_time = validate_positive_float(ls_time, "ls_time")
    x = np.array(validate_time_x(x, times), dtype=float)
    time_factor = ls / ls_time
    x[:, -1] = x[:, -1] * time_factor
    landmarks = compute_landmarks(x, n_landmarks=n_landmarks, random_state=random_state)
    if landmarks is not None:
        landmarks = np.array(landmarks, dtype=float)
        landmarks[:, -1] = landmarks[:, -1] / time_factor
    return landmarks


def compute_distances(x, k, seed=DEFAULT_RANDOM_SEED):
    """Distances to the k nearest neighbours (parameters.py:352-404).

    The reference asks pynndescent (approximate, numba-seeded) for k+1 neighbours; this uses
    scikit-learn's exact search, which is what the reference's known-answer tests
    (tests/test_parameters.py:244-268) expect and what makes neighbour selection identical on
    both sides of a parity check."""
    from sklearn.neighbors import NearestNeighbors

    x = ensure_2d(validate_array(x, "x"))
    n_samples = x.shape[0]
    if n_samples == 0:
        message = "Input data x is empty."
        logger.error(message)
        raise ValueError(message)
    validate_k(k, n_samples)
    dist, _ = NearestNeighbors(n_neighbors=k + 1).fit(x).kneighbors(x)
    return dist[:, 1:]


def compute_nn_distances(x, seed=DEFAULT_RANDOM_SEED):
    """Nearest-neighbour distance of every cell (parameters.py:407-433).

    The O(N^2 D) search runs on the device (``mb_nn_distances``: brute force over the distance tiles of K1 with a
    running-minimum epilogue, the selected pair's distance recomputed as ``sqrt(sum (x - y)^2)``; with a communicator
    every rank searches its row block against all cells).  It is exact, which is what the reference's known-answer
    tests expect from its approximate pynndescent call (``tests/test_parameters.py:244-268``), and it agrees with
    scikit-learn's exact search to the last bits (``tests/test_kernels_parity.py``, ``tests/test_golden_reference.py``)."""
    x = ensure_2d(validate_array(x, "x"))
    n_samples = x.shape[0]
    if n_samples == 0:
        message = "Input data x is empty."
        logger.error(message)
        raise ValueError(message)
    validate_k(1, n_samples)
    return get_backend().nn_distances(x)


def _get_target_cell_count(normalize, time, av_cells_per_tp, unique_times):
    if isinstance(normalize, bool):
        return av_cells_per_tp
    if isinstance(normalize, dict):
        return normalize[time.item()]
    return normalize[unique_times.tolist().index(time)]


def compute_nn_distances_within_time_points(x, times=None, d=None, normalize=False):
    """NN distances computed inside each time point (parameters.py:444-531)."""
    x = validate_time_x(x, times)
    unique_times = np.unique(x[:, -1])
    nn_distances = np.empty(x.shape[0])
    n_cells = x.shape[0]
    av_cells_per_tp = n_cells / len(unique_times)
    validate_normalize_parameter(normalize, unique_times)
    normalizing = normalize is not False and normalize is not None
    if normalizing:
        d = validate_float_or_iterable_numerical(d, "d", optional=False, positive=True)
        if np.ndim(d) > 0 and len(d) != x.shape[0]:
            raise ValueError(
                f"If `d` (length={len(d):,}) is a vector then it needs to have one value "
                f"per cell in x (x.shape[0]={n_cells:,})."
            )
        logger.info(
            "Normalizing nearest neighbor distances correcting sampling bias for "
            f"{len(unique_times):,} different time points."
        )
    for time in unique_times:
        mask = x[:, -1] == time
        n_samples = int(np.sum(mask))
        if n_samples < 2:
            raise ValueError(
                f"Insufficient data: Only {n_samples} sample(s) found at time point {time}. "
                "Nearest neighbors cannot be computed with less than two samples per time point. "
                "Please confirm if you have provided the correct time axis. "
                "If the time points indeed have very few samples, consider aggregating nearby "
                "time points for better results, or you may specify `nn_distances` manually."
            )
        at_time = compute_nn_distances(x[mask, :-1])
        if normalizing:
            target = _get_target_cell_count(normalize, time, av_cells_per_tp, unique_times)
            at_time = (n_samples / target) ** (1 / d if np.ndim(d) == 0 else 1 / d[mask]) * at_time
        nn_distances[mask] = at_time
    return nn_distances


def compute_d(x):
    """Embedding dimensionality (parameters.py:534-542)."""
    return 1 if len(np.shape(x)) < 2 else np.shape(x)[1]


def compute_d_factal(x, k=10, n=500, seed=432):
    """Average local fractal dimension (parameters.py:545-583).  Outside this package's path (SURVEY.md §2:
    `d_method="fractal"` draws its sample with jax.random and belongs to the dimensionality estimator's family);
    the name exists so that code importing it from ``mellon.parameters`` loads, and says so when called."""
    raise NotImplementedError(
        "compute_d_factal (d_method='fractal') is outside mellon_b200's path; pass d= explicitly or use "
        "d_method='embedding'."
    )


def _quantile_linear(a, q):
    """``np.quantile(a, q)`` (default linear interpolation) of a 1-D float array, bit for bit, from ONE selection
    instead of NumPy's three-pivot partition (order statistics k and k + 1 plus the NaN sentinel at -1): 2.7 ms instead
    of 12.7 ms at 1e6 cells.  It is replicated host work of every rank and every fit, i.e. pure Amdahl overhead of the
    multi-GPU step."""
    a = np.asarray(a, dtype=float)
    n = a.shape[0]
    if a.ndim != 1 or n < 4096:
        return np.quantile(a, q)
    virtual = q * (n - 1)
    k = int(np.floor(virtual))
    if k + 1 >= n:
        return np.quantile(a, q)
    gamma = virtual - k
    head = np.partition(a, k + 1)[: k + 2]
    lo, hi = head[: k + 1].max(), head[k + 1]
    if np.isnan(a.max()):                      # NumPy: any NaN makes the quantile NaN
        return np.float64(np.nan)
    diff = hi - lo                             # numpy.lib._function_base_impl._lerp
    return np.float64(hi - diff * (1 - gamma) if gamma >= 0.5 else lo + diff * gamma)


def compute_mu(nn_distances, d):
    """1st percentile of the NN maximum-likelihood log density, minus 10 (parameters.py:586-599)."""
    return float(_quantile_linear(mle(nn_distances, d), 0.01)) - 10


def compute_ls(nn_distances):
    """exp(mean(log nn) + 3) (parameters.py:602-613)."""
    return float(np.exp(np.log(np.asarray(nn_distances, dtype=float)).mean() + 3.0))


def compute_cov_func(cov_func_curry, ls, ls_time=None):
    """Instantiate the covariance; with ``ls_time`` the product state x time kernel
    (parameters.py:616-645)."""
    if ls_time is not None:
        return cov_func_curry(ls=ls, active_dims=slice(None, -1)) * cov_func_curry(ls=ls_time, active_dims=-1)
    return cov_func_curry(ls=ls)


def compute_Lp(x, cov_func, gp_type=None, landmarks=None, sigma=DEFAULT_SIGMA, jitter=DEFAULT_JITTER):
    """Cholesky factor of the (landmark) covariance, or None for Nystroem types
    (parameters.py:648-714)."""
    x = ensure_2d(x)
    n_samples = x.shape[0]
    if landmarks is None:
        n_landmarks, landmarks = n_samples, x
    else:
        landmarks = ensure_2d(landmarks)
        n_landmarks = landmarks.shape[0]
    gp_type = GaussianProcessType.from_string(gp_type, optional=True)
    if gp_type is None:
        gp_type = compute_gp_type(n_landmarks, 1.0, n_samples)
    if gp_type in _NYSTROEM:
        return None
    if gp_type == GaussianProcessType.FULL:
        logger.info("Computing Lp.")
        return _full_rank(x, cov_func, sigma=sigma, jitter=jitter)
    if gp_type in _LANDMARK_CHOLESKY:
        return _full_rank(landmarks, cov_func, sigma=sigma, jitter=jitter)
    message = f"Unknown Gaussian Process type {gp_type}."
    logger.error(message)
    raise ValueError(message)


def validate_compute_L_input(x, cov_func, gp_type, landmarks, Lp, rank, sigma, jitter):
    """Argument normalisation for compute_L (parameters.py:717-780)."""
    jitter = validate_positive_float(jitter, "jitter")
    rank = validate_float_or_int(rank, "rank", optional=True)
    n_samples = x.shape[0]
    n_landmarks = n_samples if landmarks is None else landmarks.shape[0]
    gp_type = GaussianProcessType.from_string(gp_type, optional=True)
    if rank is None:
        rank = compute_rank(gp_type)
    if gp_type is None:
        gp_type = compute_gp_type(n_landmarks, rank, n_samples)
    validate_params(rank, gp_type, n_samples, n_landmarks, landmarks)
    if gp_type == GaussianProcessType.FULL and Lp is not None and tuple(Lp.shape) != (n_samples, n_samples):
        message = f" Wrong shape of Lp {Lp.shape} for {gp_type} and {n_samples:,} samples."
        logger.error(message)
        raise ValueError(message)
    if gp_type in _LANDMARK_CHOLESKY and Lp is not None and tuple(Lp.shape) != (n_landmarks, n_landmarks):
        message = f" Wrong shape of Lp {Lp.shape} for {gp_type} and {n_landmarks:,} landmarks."
        logger.error(message)
        raise ValueError(message)
    x = ensure_2d(x)
    if landmarks is not None:
        landmarks = ensure_2d(landmarks)
    return x, landmarks, n_landmarks, n_samples, gp_type, rank


def compute_L(x, cov_func, gp_type=None, landmarks=None, Lp=None, rank=None, sigma=DEFAULT_SIGMA,
              jitter=DEFAULT_JITTER):
    """The N x r factor ``L`` for every gp_type (parameters.py:783-874)."""
    x, landmarks, n_landmarks, n_samples, gp_type, rank = validate_compute_L_input(
        x, cov_func, gp_type, landmarks, Lp, rank, sigma, jitter
    )
    if gp_type == GaussianProcessType.FULL:
        if Lp is None:
            return _full_rank(x, cov_func, sigma=sigma, jitter=jitter)
        return Lp
    if gp_type == GaussianProcessType.FULL_NYSTROEM:
        return _full_decomposition_low_rank(x, cov_func, rank=rank, sigma=sigma, jitter=jitter)
    if gp_type in _LANDMARK_CHOLESKY:
        return _standard_low_rank(x, cov_func, landmarks, Lp=Lp, sigma=sigma, jitter=jitter)
    if gp_type == GaussianProcessType.SPARSE_NYSTROEM:
        return _modified_low_rank(x, cov_func, landmarks, rank=rank, sigma=sigma, jitter=jitter)


def compute_initial_value(nn_distances, d, mu, L):
    """Ridge start ``(L^T L + I)^-1 L^T (mle - mu)`` (parameters.py:877-896).

    The reference calls sklearn ``Ridge(alpha=1, fit_intercept=False)``, whose Cholesky solver
    forms exactly these normal equations (dual form when r > N); here the Gram is K4 on the
    tensor pipe with the all-reduce over cell shards, followed by K2 and two triangular solves."""
    be = get_backend()
    target = mle(nn_distances, d) - mu
    Ld = L if isinstance(L, DeviceArray) else be.upload(np.asarray(L, dtype=float), sharded=True)
    n, r = Ld.shape
    if r <= n:
        return be.ridge_init(Ld, target)
    # dual form (r > N): z0 = L^T (L L^T + I)^-1 t
    G = be.gemm(Ld, Ld, trans_b=True)
    be.add_diag(G, 1.0)
    if be.potrf(G) > 0:
        raise ValueError("L L^T + I is not positive definite.")
    v = be.tri_solve(G, be.tri_solve(G, target), trans=True)
    return be.gemv_t(Ld, v)


def compute_average_cell_count(x, normalize):
    """Average cells per time point (parameters.py:927-969)."""
    n_cells = x.shape[0]
    n_unique_times = np.unique(np.asarray(x)[:, -1]).shape[0]
    if normalize is None or isinstance(normalize, bool):
        return n_cells / n_unique_times
    if isinstance(normalize, dict):
        return sum(normalize.values()) / n_unique_times
    if isinstance(normalize, (list, np.ndarray)):
        return np.sum(np.asarray(normalize)) / len(normalize)
    raise ValueError(f"Unrecognized type for 'normalize': {type(normalize)}")
