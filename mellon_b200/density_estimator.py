"""``DensityEstimator`` — the drop-in for ``mellon.DensityEstimator`` (``mellon/density_estimator.py``).

Same constructor, attributes and methods (``prepare_inference`` / ``run_inference`` /
``process_inference`` / ``fit`` / ``predict`` / ``fit_predict``).  The pipeline is the
reference's; each numeric step lands on the GPU:

=====================  ===========================================  ==================
step                   reference                                    device kernels
=====================  ===========================================  ==================
``Lp``                 parameters.py:648-714, decomposition.py:79   K1 + K2
``L``                  parameters.py:783-874, decomposition.py:174  K1 + K3 (+K4, eigh)
``initial_value``      parameters.py:877-896 (sklearn Ridge)        K4 + all-reduce + K2
L-BFGS-B evaluations   inference.py:167-192, 272-288                K5 + all-reduce
Laplace std            inference.py:291-338                         K6 + all-reduce
``log_density_x``      inference.py:341-354                         row-dot pass
``predict``            conditional.py + base_predictor.py:180-257   K7
=====================  ===========================================  ==================
"""

from __future__ import annotations

import logging

from .base_model import DEFAULT_COV_FUNC, BaseEstimator
from .inference import (
    DEFAULT_INIT_LEARN_RATE,
    DEFAULT_JIT,
    DEFAULT_N_ITER,
    DEFAULT_OPTIMIZER,
    compute_conditional,
    compute_log_density_x,
    compute_loss_func,
    compute_transform,
)
from .parameters import DEFAULT_RANDOM_SEED, compute_d, compute_initial_value, compute_mu
from .util import DEFAULT_JITTER, object_html
from .validation import validate_array, validate_string

DEFAULT_D_METHOD = "embedding"

logger = logging.getLogger("mellon")


class DensityEstimator(BaseEstimator):
    """Non-parametric density estimator: a GP prior on the log density and the
    nearest-neighbour-distance likelihood (``density_estimator.py:38-581``)."""

    def __init__(self, cov_func_curry=DEFAULT_COV_FUNC, n_landmarks=None, rank=None, gp_type=None,
                 d_method=DEFAULT_D_METHOD, jitter=DEFAULT_JITTER, optimizer=DEFAULT_OPTIMIZER,
                 n_iter=DEFAULT_N_ITER, init_learn_rate=DEFAULT_INIT_LEARN_RATE, landmarks=None,
                 nn_distances=None, d=None, mu=None, ls=None, ls_factor=1, cov_func=None, Lp=None, L=None,
                 initial_value=None, predictor_with_uncertainty=False, jit=DEFAULT_JIT, check_rank=None,
                 random_state=DEFAULT_RANDOM_SEED):
        super().__init__(
            cov_func_curry=cov_func_curry, n_landmarks=n_landmarks, rank=rank, jitter=jitter, gp_type=gp_type,
            optimizer=optimizer, n_iter=n_iter, init_learn_rate=init_learn_rate, landmarks=landmarks,
            nn_distances=nn_distances, d=d, mu=mu, ls=ls, ls_factor=ls_factor, cov_func=cov_func, Lp=Lp, L=L,
            initial_value=initial_value, predictor_with_uncertainty=predictor_with_uncertainty, jit=jit,
            check_rank=check_rank, random_state=random_state,
        )
        if d is not None:
            self.d_method = "manual"
            logger.info(f"Explicitly provided d={d}, setting d_method to 'manual'.")
        else:
            self.d_method = validate_string(d_method, "d_method", choices={"fractal", "embedding", "manual"})
        self.transform = None
        self.loss_func = None
        self.opt_state = None
        self.losses = None
        self.pre_transformation = None
        self.pre_transformation_std = None
        self.log_density_x = None
        self.log_density_func = None

    def _repr_html_(self):
        rows = {
            "Jitter": self.jitter,
            "Mean (μ)": self.mu or "Not Set",
            "Length Scale (ls)": self.ls or "Not Set",
            "Length-Scale Factor": self.ls_factor,
            "Dimensionality (d)": self.d if self.d is not None else "Not Set",
            "Nearest Neighbor Distances": self.nn_distances,
            "Landmarks": self.landmarks,
            "L": self.L,
            "Lp": self.Lp,
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
        """``density_estimator.py:311-333``"""
        if self.d_method == "fractal":
            raise NotImplementedError(
                "d_method='fractal' (mellon.parameters.compute_d_factal) is outside mellon_b200's path; "
                "compute it with mellon and pass d=... explicitly."
            )
        if self.d_method == "manual":
            d = self.d
            logger.info(f"Using manually set d={d}.")
        else:
            d = compute_d(self.x)
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
        k = self.initial_value.shape[0]
        return compute_loss_func(self.nn_distances, self.d, self.transform, k)

    def _set_log_density_x(self):
        self.log_density_x = compute_log_density_x(self.pre_transformation, self.transform)

    def _set_log_density_func(self):
        """Build the predictor (``density_estimator.py:370-402``)."""
        logger.info("Computing predictive function.")
        log_density_func = compute_conditional(
            self.x, self.landmarks, self.pre_transformation, self.pre_transformation_std, self.log_density_x,
            self.mu, self.cov_func, self.L, self.Lp, sigma=None, jitter=self.jitter, y_is_mean=True,
            with_uncertainty=self.predictor_with_uncertainty,
        )
        log_density_func.n_obs = self.x.shape[0]
        log_density_func.d = self.d
        log_density_func.d_method = self.d_method
        self.log_density_func = log_density_func

    # -- public pipeline ----------------------------------------------------------------------
    def prepare_inference(self, x):
        """Compute every missing parameter; returns ``(loss_func, initial_value)``
        (``density_estimator.py:404-444``)."""
        if x is None:
            x = self.x
            if self.x is None:
                self._fail("Required argument x is missing and self.x has not been set.")
        elif self.x is not None and self.x is not x:
            self._fail("self.x has been set already, but is not equal to the argument x.")
        self.set_x(x)
        self._prepare_attribute("n_landmarks")
        self._prepare_attribute("rank")
        self._prepare_attribute("gp_type")
        self.validate_parameter()
        self._prepare_attribute("nn_distances")
        self._prepare_attribute("d")
        self._prepare_attribute("mu")
        self._prepare_attribute("ls")
        self._prepare_attribute("cov_func")
        self._prepare_attribute("landmarks")
        self._prepare_attribute("Lp")
        self._prepare_attribute("L")
        self._prepare_attribute("initial_value")
        self._prepare_attribute("transform")
        self._prepare_attribute("loss_func")
        return self.loss_func, self.initial_value

    def run_inference(self, loss_func=None, initial_value=None, optimizer=None):
        """Minimise the loss; returns the optimal pre-transformation (``:446-469``)."""
        if loss_func is not None:
            self.loss_func = loss_func
        if initial_value is not None:
            self.initial_value = initial_value
        if optimizer is not None:
            self.optimizer = optimizer
        self._run_inference()
        return self.pre_transformation

    def process_inference(self, pre_transformation=None, build_predict=True):
        """Turn the optimum into log densities (and the predictor) (``:471-492``)."""
        if pre_transformation is not None:
            self.pre_transformation = validate_array(pre_transformation, "pre_transformation")
        self._set_log_density_x()
        if build_predict:
            self._set_log_density_func()
        return self.log_density_x

    def fit(self, x=None, build_predict=True):
        """Fit the model from end to end (``:494-516``)."""
        self.prepare_inference(x)
        self.run_inference()
        self.process_inference(build_predict=build_predict)
        return self

    @property
    def predict(self):
        """The log-density predictor, built on first access (``:518-540``)."""
        if self.log_density_func is None:
            self._set_log_density_func()
        return self.log_density_func

    def fit_predict(self, x=None, build_predict=False):
        """Fit and return the log density at the training points (``:542-581``)."""
        if self.x is not None and x is not None and self.x is not x:
            self._fail("self.x has been set already, but is not equal to the argument x.")
        if self.x is None and x is None:
            self._fail("Required argument x is missing and self.x has not been set.")
        x = self.x if x is None else validate_array(x, "x")
        self.fit(x, build_predict=build_predict)
        return self.log_density_x
