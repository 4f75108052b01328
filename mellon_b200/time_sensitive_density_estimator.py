"""``TimeSensitiveDensityEstimator`` — drop-in for ``mellon.TimeSensitiveDensityEstimator``
(``mellon/time_sensitive_density_estimator.py``).

Time is the last column of ``x``; the covariance is the product of a state kernel on
``active_dims=slice(None, -1)`` and a time kernel on ``active_dims=-1``
(``parameters.py:641-644``), which the device evaluates as ONE two-leaf covariance program inside
the same fused kernel (K1 / K7) as the plain estimators.
"""

from __future__ import annotations

import logging

from .base_model import DEFAULT_COV_FUNC, BaseEstimator
from .compute_ls_time import compute_ls_time
from .inference import (
    DEFAULT_INIT_LEARN_RATE,
    DEFAULT_JIT,
    DEFAULT_N_ITER,
    DEFAULT_OPTIMIZER,
    compute_conditional_times,
    compute_log_density_x,
    compute_loss_func,
    compute_transform,
)
from .parameters import (
    DEFAULT_RANDOM_SEED,
    compute_average_cell_count,
    compute_cov_func,
    compute_d,
    compute_initial_value,
    compute_landmarks_rescale_time,
    compute_ls,
    compute_mu,
    compute_nn_distances_within_time_points,
)
from .util import DEFAULT_JITTER, object_html, object_str
from .validation import (
    validate_array,
    validate_nn_distances,
    validate_positive_float,
    validate_string,
    validate_time_x,
)

DEFAULT_D_METHOD = "embedding"

logger = logging.getLogger("mellon")


class TimeSensitiveDensityEstimator(BaseEstimator):
    """Density estimator over (state, time) (``time_sensitive_density_estimator.py:45-796``)."""

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
        if d is not None:
            self.d_method = "manual"
            logger.info(f"Explicitly provided d={d}, setting d_method to 'manual'.")
        else:
            self.d_method = validate_string(d_method, "d_method", choices={"fractal", "embedding", "manual"})
        self.ls_time = validate_positive_float(ls_time, "ls_time", optional=True)
        self.ls_time_factor = validate_positive_float(ls_time_factor, "ls_time_factor")
        self._save_intermediate_ls_times = _save_intermediate_ls_times
        self.normalize_per_time_point = normalize_per_time_point
        self.transform = None
        self.loss_func = None
        self.opt_state = None
        self.losses = None
        self.pre_transformation = None
        self.pre_transformation_std = None
        self.log_density_x = None
        self.log_density_func = None

    def __repr__(self):
        head = super().__repr__()[:-2]
        return (
            head
            + f"\n    ls_time={self.ls_time},"
            + f"\n    ls_time_factor={self.ls_time_factor},"
            + f"\n    normalize_per_time_point={object_str(self.normalize_per_time_point)},"
            + f"\n    density_estimator_kwargs={self.density_estimator_kwargs},"
            + "\n)"
        )

    def _repr_html_(self):
        rows = {
            "Jitter": self.jitter,
            "Mean (μ)": self.mu or "Not Set",
            "Length Scale (ls)": self.ls or "Not Set",
            "Time Length Scale (ls_time)": self.ls_time or "Not Set",
            "Dimensionality (d)": self.d if self.d is not None else "Not Set",
            "Nearest Neighbor Distances": self.nn_distances,
            "Landmarks": self.landmarks,
        }
        table = "".join(f"<tr><td>{k}</td><td>{object_html(v)}</td></tr>" for k, v in rows.items())
        status = "Available" if self.log_density_func else "Not Yet Computed"
        return (
            f"<div><h3>{self.__class__.__name__}</h3>"
            f"<p>gp_type={self.gp_type}, optimizer={self.optimizer}, covariance={self.cov_func!r}</p>"
            f"<table><tr><th>Parameter</th><th>Value</th></tr>{table}</table>"
            f"<p><strong>Predictor:</strong> {status}</p></div>"
        )

    # -- lazy attribute pipeline ------------------------------------------------------------
    def _compute_d(self):
        """``time_sensitive_density_estimator.py:423-447`` — d of the state columns."""
        x = self.x[:, :-1]
        if self.d_method == "fractal":
            raise NotImplementedError(
                "d_method='fractal' (mellon.parameters.compute_d_factal) is outside mellon_b200's path; "
                "compute it with mellon and pass d=... explicitly."
            )
        if self.d_method == "manual":
            d = self.d
            logger.info(f"Using manually set d={d}.")
        else:
            d = compute_d(x)
            logger.info(
                f"Using embedding dimensionality d={d}. "
                'Use d_method="fractal" to enable effective density normalization.'
            )
        if d > 50:
            raise ValueError(
                "The detected dimensionality of the data is over 50, which is likely to cause numerical "
                "instability issues. Consider running a dimensionality reduction algorithm, or if this "
                f"number of dimensions is intended, explicitly pass d={self.d} as a parameter."
            )
        return d

    def _compute_mu(self):
        return compute_mu(self.nn_distances, self.d)

    def _compute_initial_value(self):
        return compute_initial_value(self.nn_distances, self.d, self.mu, self.L)

    def _compute_transform(self):
        return compute_transform(self.mu, self.L)

    def _compute_loss_func(self):
        return compute_loss_func(self.nn_distances, self.d, self.transform, self.initial_value.shape[0])

    def _compute_nn_distances(self):
        logger.info("Computing nearest neighbor distances within time points.")
        nn = compute_nn_distances_within_time_points(self.x, d=self.d, normalize=self.normalize_per_time_point)
        return validate_nn_distances(nn)

    def _compute_ls(self):
        nn_distances = self.nn_distances
        normalized = self.normalize_per_time_point
        if normalized is not False and normalized is not None:
            logger.info("Computing non-normalized nn_distances for length scale heuristic.")
            nn_distances = compute_nn_distances_within_time_points(self.x, normalize=False)
        return compute_ls(nn_distances) * self.ls_factor

    def _compute_ls_time(self):
        """One density fit per time point (``:503-536``); pass ``ls_time`` to skip it."""
        kwargs = {
            "cov_func_curry": self.cov_func_curry,
            "d_method": self.d_method,
            "d": self.d,
            "optimizer": self.optimizer,
            "ls": self.ls,
            "ls_factor": self.ls_factor,
            "jit": self.jit,
            "mu": self.mu,
        }
        kwargs.update(self.density_estimator_kwargs)
        logger.info(
            "Initiating density computation for each time point to estimate the 'ls_time' parameter. "
            "You can directly specify 'ls_time' to bypass this computation-intensive step."
        )
        ls = compute_ls_time(self.nn_distances, self.x, self.cov_func_curry,
                             return_data=self._save_intermediate_ls_times, density_estimator_kwargs=kwargs)
        if self._save_intermediate_ls_times:
            logger.info("Storing `self.densities`, `self.predictors`, and `self.numeric_stages`.")
            ls, self.densities, self.predictors, self.numeric_stages = ls
        return ls * self.ls_time_factor

    def _compute_landmarks(self):
        n_samples, n_landmarks = self.x.shape[0], self.n_landmarks
        if n_samples > 100 * n_landmarks and n_samples > 1e6:
            logger.info(
                f"Large number of {n_samples:,} cells and small number of {n_landmarks:,} landmarks. "
                "Consider computing k-means on a subset of cells and passing the results as "
                "'landmarks' to speed up the process."
            )
        return compute_landmarks_rescale_time(self.x, self.ls, self.ls_time, n_landmarks=n_landmarks,
                                              random_state=self._seed())

    def _compute_cov_func(self):
        cov_func = compute_cov_func(self.cov_func_curry, self.ls, self.ls_time)
        logger.info("Using covariance function %s.", str(cov_func))
        return cov_func

    def _set_log_density_x(self):
        self.log_density_x = compute_log_density_x(self.pre_transformation, self.transform)

    def _set_log_density_func(self):
        """``:573-606``"""
        logger.info("Computing predictive function.")
        log_density_func = compute_conditional_times(
            self.x, self.landmarks, self.pre_transformation, self.pre_transformation_std, self.log_density_x,
            self.mu, self.cov_func, self.L, self.Lp, sigma=None, jitter=self.jitter, y_is_mean=True,
            with_uncertainty=self.predictor_with_uncertainty,
        )
        log_density_func.n_obs = compute_average_cell_count(self.x, self.normalize_per_time_point)
        log_density_func.d = self.d
        log_density_func.d_method = self.d_method
        self.log_density_func = log_density_func

    # -- public pipeline ----------------------------------------------------------------------
    def prepare_inference(self, x, times=None):
        """``:608-665`` — note the order: ``d`` before ``nn_distances``, ``ls_time`` after ``ls``."""
        if x is None:
            x = self.x
            if self.x is None:
                self._fail("Required argument x is missing and self.x has not been set.")
        else:
            x = validate_time_x(x, times)
            if self.x is not None and self.x is not x:
                self._fail("self.x has been set already, but is not equal to the argument x.")
        self.set_x(x)
        self._prepare_attribute("n_landmarks")
        self._prepare_attribute("rank")
        self._prepare_attribute("gp_type")
        self.validate_parameter()
        self._prepare_attribute("d")
        self._prepare_attribute("nn_distances")
        self._prepare_attribute("mu")
        self._prepare_attribute("ls")
        self._prepare_attribute("ls_time")
        self._prepare_attribute("cov_func")
        self._prepare_attribute("landmarks")
        self._prepare_attribute("Lp")
        self._prepare_attribute("L")
        self._prepare_attribute("initial_value")
        self._prepare_attribute("transform")
        self._prepare_attribute("loss_func")
        return self.loss_func, self.initial_value

    def run_inference(self, loss_func=None, initial_value=None, optimizer=None):
        if loss_func is not None:
            self.loss_func = loss_func
        if initial_value is not None:
            self.initial_value = initial_value
        if optimizer is not None:
            self.optimizer = optimizer
        self._run_inference()
        return self.pre_transformation

    def process_inference(self, pre_transformation=None, build_predict=True):
        if pre_transformation is not None:
            self.pre_transformation = validate_array(pre_transformation, "pre_transformation")
        self._set_log_density_x()
        if build_predict:
            self._set_log_density_func()
        return self.log_density_x

    def fit(self, x=None, times=None, build_predict=True):
        self.prepare_inference(x, times)
        self.run_inference()
        self.process_inference(build_predict=build_predict)
        return self

    @property
    def predict(self):
        if self.log_density_func is None:
            self._set_log_density_func()
        return self.log_density_func

    def fit_predict(self, x=None, times=None, build_predict=False):
        """``:772-796``"""
        if x is not None:
            x = validate_time_x(x, times)
        if self.x is not None and x is not None and self.x is not x:
            self._fail("self.x has been set already, but is not equal to the argument x.")
        if self.x is None and x is None:
            self._fail("Required argument x is missing and self.x has not been set.")
        if x is None:
            x = self.x
        self.fit(x, build_predict=build_predict)
        return self.log_density_x
