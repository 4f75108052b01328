"""What ``DensityEstimator`` and ``TimeSensitiveDensityEstimator`` have in common.

The reference spells the density pipeline out twice (``mellon/density_estimator.py:311-540`` and
``mellon/time_sensitive_density_estimator.py:423-770``): the dimensionality rule, ``mu`` / ``initial_value`` / ``transform``
/ ``loss_func`` from the nearest-neighbour distances, the inference triple ``run_inference`` / ``process_inference`` /
``predict``, and the bookkeeping on the predictor.  Here it lives once; the two estimators add their constructor, the order
of their lazy attributes (``PIPELINE``) and their predictor.
"""

from __future__ import annotations

import logging

from .base_model import BaseEstimator
from .inference import compute_log_density_x, compute_loss_func, compute_transform
from .parameters import compute_d, compute_initial_value, compute_mu
from .util import object_html
from .validation import validate_array, validate_string

DEFAULT_D_METHOD = "embedding"
D_METHODS = {"fractal", "embedding", "manual"}
MAX_D = 50

logger = logging.getLogger("mellon")


class DensityPipeline(BaseEstimator):
    """Shared part of the two density estimators."""

    #: attributes a fit produces; ``None`` until then
    RESULTS = ("transform", "loss_func", "opt_state", "losses", "pre_transformation", "pre_transformation_std",
               "log_density_x", "log_density_func")
    #: lazy attributes in the order ``prepare_inference`` fills them (set by the subclasses)
    PIPELINE = ()

    def _init_density(self, d, d_method):
        """``d`` given explicitly wins over ``d_method`` (``density_estimator.py:207-214``)."""
        if d is None:
            self.d_method = validate_string(d_method, "d_method", choices=D_METHODS)
        else:
            logger.info(f"Explicitly provided d={d}, setting d_method to 'manual'.")
            self.d_method = "manual"
        for name in self.RESULTS:
            setattr(self, name, None)

    def _html(self, rows):
        """The notebook card (``_repr_html_``): one table row per entry of ``rows``."""
        table = "".join(f"<tr><td>{k}</td><td>{object_html(v)}</td></tr>" for k, v in rows.items())
        status = "Available" if self.log_density_func else "Not Yet Computed"
        return (
            f"<div><h3>{self.__class__.__name__}</h3>"
            f"<p>gp_type={self.gp_type}, optimizer={self.optimizer}, covariance={self.cov_func!r}</p>"
            f"<table><tr><th>Parameter</th><th>Value</th></tr>{table}</table>"
            f"<p><strong>Predictor:</strong> {status}</p></div>"
        )

    # -- lazy attributes ------------------------------------------------------------------------
    def _state(self):
        """The columns the dimensionality is computed from (all of ``x``; without the time column for the time-sensitive
        estimator)."""
        return self.x

    def _compute_d(self):
        """``density_estimator.py:311-333`` / ``time_sensitive_density_estimator.py:423-447``"""
        method = self.d_method
        if method == "fractal":
            raise NotImplementedError(
                "d_method='fractal' (mellon.parameters.compute_d_factal) is outside mellon_b200's path; "
                "compute it with mellon and pass d=... explicitly."
            )
        if method == "manual":
            d = self.d
            logger.info(f"Using manually set d={d}.")
        else:
            d = compute_d(self._state())
            logger.info(f"Using embedding dimensionality d={d}. "
                        'Use d_method="fractal" to enable effective density normalization.')
        if d > MAX_D:
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

    def _set_log_density_x(self):
        self.log_density_x = compute_log_density_x(self.pre_transformation, self.transform)

    def _build_predictor(self, conditional, n_obs):
        """Shared tail of ``_set_log_density_func`` (``density_estimator.py:370-402``): ``conditional`` is
        ``compute_conditional`` or ``compute_conditional_times``."""
        logger.info("Computing predictive function.")
        func = conditional(self.x, self.landmarks, self.pre_transformation, self.pre_transformation_std,
                           self.log_density_x, self.mu, self.cov_func, self.L, self.Lp, sigma=None, jitter=self.jitter,
                           y_is_mean=True, with_uncertainty=self.predictor_with_uncertainty)
        func.n_obs, func.d, func.d_method = n_obs, self.d, self.d_method
        self.log_density_func = func

    # -- public pipeline --------------------------------------------------------------------------
    def _claim_x(self, x, validate=None, validate_first=True):
        """The ``x`` rule of ``prepare_inference`` / ``fit_predict`` (``density_estimator.py:425-431, 566-576``): a given
        ``x`` must be THE array already set, a missing one falls back to ``self.x``.  The time-sensitive estimator
        validates (appends the time column) before the identity check, ``DensityEstimator.fit_predict`` after it."""
        if x is not None and validate is not None and validate_first:
            x = validate(x)
        if x is None and self.x is None:
            self._fail("Required argument x is missing and self.x has not been set.")
        if x is not None and self.x is not None and self.x is not x:
            self._fail("self.x has been set already, but is not equal to the argument x.")
        if x is None:
            return self.x
        return validate(x) if validate is not None and not validate_first else x

    def _prepare_pipeline(self, x):
        self.set_x(x)
        for name in ("n_landmarks", "rank", "gp_type"):
            self._prepare_attribute(name)
        self.validate_parameter()
        for name in self.PIPELINE:
            self._prepare_attribute(name)
        return self.loss_func, self.initial_value

    def run_inference(self, loss_func=None, initial_value=None, optimizer=None):
        """Minimise the loss; returns the optimal pre-transformation (``density_estimator.py:446-469``)."""
        for name, given in (("loss_func", loss_func), ("initial_value", initial_value), ("optimizer", optimizer)):
            if given is not None:
                setattr(self, name, given)
        self._run_inference()
        return self.pre_transformation

    def process_inference(self, pre_transformation=None, build_predict=True):
        """Turn the optimum into log densities (and the predictor) (``density_estimator.py:471-492``)."""
        if pre_transformation is not None:
            self.pre_transformation = validate_array(pre_transformation, "pre_transformation")
        self._set_log_density_x()
        if build_predict:
            self._set_log_density_func()
        return self.log_density_x

    @property
    def predict(self):
        """The log-density predictor, built on first access (``density_estimator.py:518-540``)."""
        if self.log_density_func is None:
            self._set_log_density_func()
        return self.log_density_func
