"""Time length-scale heuristic (``mellon/compute_ls_time.py``): one ``DensityEstimator`` fit per
time point (each of them the accelerated path), then a 1-D fit of the time kernel to the
correlation of the per-time-point densities."""

from __future__ import annotations

import logging

import numpy as np
from scipy.optimize import minimize

from .density_estimator import DensityEstimator
from .validation import validate_time_x

logger = logging.getLogger("mellon")


def compute_ls_time(nn_distances, x, cov_func_curry, times=None, warn_below=500, return_data=False,
                    density_estimator_kwargs=dict()):
    """``compute_ls_time.py:12-104``.  The 1-D objective is tiny (n_times^2 kernel values) and is
    evaluated on the host with the closed forms of the stock kernels through ``cov(...)``; the
    reference differentiates it with JAX, here SciPy's L-BFGS-B uses its own finite differences."""
    x = validate_time_x(x, times)
    times = x[:, -1]
    states = np.ascontiguousarray(x[:, :-1])
    unique_times = np.unique(times)
    n_times = len(unique_times)
    nn_distances = np.asarray(nn_distances, dtype=float)
    densities, predictors = [], []
    for i, time in enumerate(unique_times):
        mask = times == time
        n_cells = int(np.sum(mask))
        logger.info(f"[{i+1} of {n_times}] Computing density for {n_cells:,} cells at time point {time}.")
        if n_cells < warn_below:
            logger.warning(
                f"Time point {time} only has {n_cells:,} cells. "
                "This could lead to inaccurate estimation of the time length scale `ls_time`."
            )
        est = DensityEstimator(nn_distances=nn_distances[mask], **density_estimator_kwargs)
        est.fit(np.ascontiguousarray(x[mask, :-1]))
        densities.append(np.asarray(est.predict(states)))
        predictors.append(est)
    densities = np.stack(densities)
    corrs = np.corrcoef(densities)
    delta_t = np.abs(unique_times.reshape(-1, 1) - unique_times.reshape(1, -1)).reshape(-1, 1)
    origin = np.zeros((1, 1))

    def ls_loss(log_ls):
        ls = float(np.exp(np.ravel(log_ls)[0]))
        covs = np.asarray(cov_func_curry(ls)(delta_t, origin)).reshape((n_times, n_times))
        return float(np.linalg.norm(covs - corrs))

    opt = minimize(ls_loss, np.zeros(1), method="L-BFGS-B", options={"maxiter": 500})
    ls = float(np.exp(opt.x[0]))
    if return_data:
        return ls, densities, predictors, unique_times
    return ls
