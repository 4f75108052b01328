"""``TimeSensitiveDensityEstimator`` — drop-in for ``mellon.TimeSensitiveDensityEstimator``
(``mellon/time_sensitive_density_estimator.py``).

Time is the last column of ``x``; the covariance is the product of a state kernel on
``active_dims=slice(None, -1)`` and a time kernel on ``active_dims=-1``
(``parameters.py:641-644``), which the device evaluates as ONE two-leaf covariance program inside
the same fused kernel (K1 / K7) as the plain estimators.  The steps shared with ``DensityEstimator`` are in
:mod:`mellon_b200.density_pipeline`.
"""

from __future__ import annotations

import logging

from . import parameters as P
from .base_model import DEFAULT_COV_FUNC
from .compute_ls_time import compute_ls_time
from .density_pipeline import DEFAULT_D_METHOD, DensityPipeline
from .inference import (DEFAULT_INIT_LEARN_RATE, DEFAULT_JIT, DEFAULT_N_ITER, DEFAULT_OPTIMIZER,
                        compute_conditional_times)
from .parameters import DEFAULT_RANDOM_SEED
from .util import DEFAULT_JITTER, object_str
from .validation import validate_nn_distances, validate_positive_float, validate_time_x

logger = logging.getLogger("mellon")


class TimeSensitiveDensityEstimator(DensityPipeline):
    """Density estimator over (state, time) (``time_sensitive_density_estimator.py:45-796``)."""

    # time_sensitive_density_estimator.py:649-664: d before the neighbour distances, ls_time after ls
    PIPELINE = ("d", "nn_distances", "mu", "ls", "ls_time", "cov_func", "landmarks", "Lp", "L", "initial_value",
                "transform", "loss_func")
    # what the per-time-point density fits of the ls_time heuristic inherit from this estimator (:514-523)
    LS_TIME_INHERITS = ("cov_func_curry", "d_method", "d", "optimizer", "ls", "ls_factor", "jit", "mu")

    def __init__(self, cov_func_curry=DEFAULT_COV_FUNC, n_landmarks=None, rank=None, gp_type=None,
                 d_method=DEFAULT_D_METHOD, jitter=DEFAULT_JITTER, optimizer=DEFAULT_OPTIMIZER,
                 n_iter=DEFAULT_N_ITER, init_learn_rate=DEFAULT_INIT_LEARN_RATE, landmarks=None,
                 nn_distances=None, normalize_per_time_point=False, d=None, mu=None, ls=None, ls_time=None,
                 ls_factor=1, ls_time_factor=1, density_estimator_kwargs=dict(), cov_func=None, Lp=None, L=None,
                 initial_value=None, predictor_with_uncertainty=False, _save_intermediate_ls_times=False,
                 jit=DEFAULT_JIT, check_rank=None, random_state=DEFAULT_RANDOM_SEED):
        super().__init__(
            cov_func_curry=cov_func_curry, n_landmarks=n_landmarks, rank=rank, jitter=jitter, gp_type=gp_type,
            optimizer=optimizer, n_iter=n_iter, init_learn_rate=init_learn_rate, landmarks=landmarks,
            nn_distances=nn_distances, d=d, mu=mu, ls=ls, ls_factor=ls_factor, cov_func=cov_func, Lp=Lp, L=L,
            initial_value=initial_value, predictor_with_uncertainty=predictor_with_uncertainty, jit=jit,
            check_rank=check_rank, random_state=random_state,
        )
        if not isinstance(density_estimator_kwargs, dict):
            raise ValueError("density_estimator_kwargs needs to be a dictionary.")
        self.density_estimator_kwargs = density_estimator_kwargs
        self._init_density(d, d_method)
        self.ls_time = validate_positive_float(ls_time, "ls_time", optional=True)
        self.ls_time_factor = validate_positive_float(ls_time_factor, "ls_time_factor")
        self._save_intermediate_ls_times = _save_intermediate_ls_times
        self.normalize_per_time_point = normalize_per_time_point

    def __repr__(self):
        extra = {"ls_time": self.ls_time, "ls_time_factor": self.ls_time_factor,
                 "normalize_per_time_point": object_str(self.normalize_per_time_point),
                 "density_estimator_kwargs": self.density_estimator_kwargs}
        return super().__repr__()[:-2] + "".join(f"\n    {k}={v}," for k, v in extra.items()) + "\n)"

    def _repr_html_(self):
        unset = "Not Set"
        return self._html({
            "Jitter": self.jitter, "Mean (μ)": self.mu or unset, "Length Scale (ls)": self.ls or unset,
            "Time Length Scale (ls_time)": self.ls_time or unset, "Dimensionality (d)": unset if self.d is None else self.d,
            "Nearest Neighbor Distances": self.nn_distances, "Landmarks": self.landmarks,
        })

    # -- lazy attributes ------------------------------------------------------------------------
    def _state(self):
        return self.x[:, :-1]

    def _time_point_distances(self, **kwargs):
        return P.compute_nn_distances_within_time_points(self.x, **kwargs)

    def _compute_nn_distances(self):
        logger.info("Computing nearest neighbor distances within time points.")
        return validate_nn_distances(self._time_point_distances(d=self.d, normalize=self.normalize_per_time_point))

    def _compute_ls(self):
        """The length-scale heuristic always sees NON-normalised distances (``:489-501``)."""
        distances = self.nn_distances
        if not any(self.normalize_per_time_point is off for off in (False, None)):   # may be an array: identity only
            logger.info("Computing non-normalized nn_distances for length scale heuristic.")
            distances = self._time_point_distances(normalize=False)
        return self.ls_factor * P.compute_ls(distances)

    def _compute_ls_time(self):
        """One density fit per time point (``:503-536``); pass ``ls_time`` to skip it."""
        kwargs = {name: getattr(self, name) for name in self.LS_TIME_INHERITS}
        kwargs.update(self.density_estimator_kwargs)
        logger.info(
            "Initiating density computation for each time point to estimate the 'ls_time' parameter. "
            "You can directly specify 'ls_time' to bypass this computation-intensive step."
        )
        keep = self._save_intermediate_ls_times
        result = compute_ls_time(self.nn_distances, self.x, self.cov_func_curry, return_data=keep,
                                 density_estimator_kwargs=kwargs)
        if keep:
            logger.info("Storing `self.densities`, `self.predictors`, and `self.numeric_stages`.")
            result, self.densities, self.predictors, self.numeric_stages = result
        return result * self.ls_time_factor

    def _compute_landmarks(self):
        n_samples, n_landmarks = self.x.shape[0], self.n_landmarks
        if n_samples > 1e6 and n_samples > 100 * n_landmarks:
            logger.info(
                f"Large number of {n_samples:,} cells and small number of {n_landmarks:,} landmarks. "
                "Consider computing k-means on a subset of cells and passing the results as "
                "'landmarks' to speed up the process."
            )
        return P.compute_landmarks_rescale_time(self.x, self.ls, self.ls_time, n_landmarks=n_landmarks,
                                                random_state=self._seed())

    def _compute_cov_func(self):
        cov_func = P.compute_cov_func(self.cov_func_curry, self.ls, self.ls_time)
        logger.info("Using covariance function %s.", str(cov_func))
        return cov_func

    def _set_log_density_func(self):
        """``:573-606``"""
        self._build_predictor(compute_conditional_times,
                              n_obs=P.compute_average_cell_count(self.x, self.normalize_per_time_point))

    # -- public pipeline --------------------------------------------------------------------------
    def prepare_inference(self, x, times=None):
        """``:608-665``"""
        return self._prepare_pipeline(self._claim_x(x, validate=lambda a: validate_time_x(a, times)))

    def fit(self, x=None, times=None, build_predict=True):
        self.prepare_inference(x, times)
        self.run_inference()
        self.process_inference(build_predict=build_predict)
        return self

    def fit_predict(self, x=None, times=None, build_predict=False):
        """``:772-796``"""
        self.fit(self._claim_x(x, validate=lambda a: validate_time_x(a, times)), build_predict=build_predict)
        return self.log_density_x
