"""``FunctionEstimator`` — the drop-in for ``mellon.FunctionEstimator`` (``mellon/function_estimator.py``).

Gaussian-process regression of observed values ``y`` (one column per output) on the cell states: there is
no optimisation, ``fit`` goes straight to the conditional of ``mellon/conditional.py`` — the full GP when
there are no landmarks, the sparse ``A A^T`` solve when there are (``inference.py:375-508`` picks them; a
``pre_transformation`` never exists on this route).  SURVEY.md §8f.3: the same device kernels as the
density path, no new ones —

=========================  ===============================================  =======================
step                       reference                                        device kernels
=========================  ===============================================  =======================
``Lp``, ``A^T``            conditional.py:513-523                           K1 + K2, K1 + K3
``A A^T``, ``A (y - mu)``  conditional.py:57-66 (``_sparse_solve``)         K4 (+ all-reduce), GEMM
full-GP factor and solves  conditional.py:233-264                           K1 + K2, TRSV
leverage                   conditional.py:312-330, 596-616, 375-400, 660    K1, K4, K2, K3, row norms
``predict`` / obs_variance conditional.py:366-373, 402-407, 651-658, 687    K7 (fused cov + mat-vec)
=========================  ===============================================  =======================
"""

from __future__ import annotations

import logging

from .base_model import DEFAULT_COV_FUNC, BaseEstimator
from . import inference as I
from . import validation as V
from .inference import DEFAULT_INIT_LEARN_RATE, DEFAULT_N_ITER, DEFAULT_OPTIMIZER
from .parameters import DEFAULT_RANDOM_SEED
from .util import DEFAULT_JITTER, GaussianProcessType, object_html

logger = logging.getLogger("mellon")

NYSTROEM_TYPES = (GaussianProcessType.FULL_NYSTROEM, GaussianProcessType.SPARSE_NYSTROEM)
# message text of the reference (function_estimator.py:171-177, 350-355, 399-403, 545-560, 598-609)
MSG_NO_NYSTROEM = ("gp_type={gp_type} but the Nyström rank reduction is "
                   "not available for the Function Estimator. "
                   "Use gp_type='cholesky' or gp_type='full' instead.")
MSG_OTHER_X = ("self.x has been set already, but is not equal to the argument x. "
               "Current landmarks might be inapropriate.")
MSG_SAMPLES = "X.shape[0] = {n:,} (n_samples) should equal y.shape[0] = {m:,}."
MSG_NDIM = ("The provided arrays, 'x' and 'Xnew', do not have the same number of dimensions. "
            "'x' is {a}-D and 'Xnew' is {b}-D. Please provide arrays with consistent dimensionality.")
MSG_FEATURES = ("The provided arrays, 'x' and 'Xnew', should have the same number of features. "
                "Got Xnew.shape[1] = {b}, but expected it to be equal to x.shape[1] = {a}. "
                "Please provide arrays with the same number of features.")
MSG_DEPRECATED = ("Deprecation Warning: FunctionEstimator's multi_fit_predict method is deprecated. "
                  "Use FunctionEstimator.fit_reodict instead.")
MSG_TRANSPOSE = ("Y.shape[0] does not equal X.shape[0] (the number of samples). "
                 "However, Y.shape[1] == X.shape[0]. Transposing Y. "
                 "This assumes the columns of Y are the samples. Please verify.")


class FunctionEstimator(BaseEstimator):
    """GP regression of function values on cell states (``function_estimator.py:29-615``)."""

    def __init__(self, cov_func_curry=DEFAULT_COV_FUNC, n_landmarks=None, gp_type=None, jitter=DEFAULT_JITTER,
                 optimizer=DEFAULT_OPTIMIZER, n_iter=DEFAULT_N_ITER, init_learn_rate=DEFAULT_INIT_LEARN_RATE,
                 landmarks=None, nn_distances=None, mu=0, ls=None, ls_factor=1, cov_func=None, sigma=0,
                 y_is_mean=False, predictor_with_uncertainty=False, obs_variance=False, jit=True,
                 random_state=DEFAULT_RANDOM_SEED):
        super().__init__(
            cov_func_curry=cov_func_curry, n_landmarks=n_landmarks, rank=1.0, jitter=jitter, gp_type=gp_type,
            landmarks=landmarks, nn_distances=nn_distances, mu=mu, ls=ls, ls_factor=ls_factor, cov_func=cov_func,
            predictor_with_uncertainty=predictor_with_uncertainty, jit=jit, random_state=random_state,
        )
        self.y_is_mean, self.obs_variance = V.validate_bool(y_is_mean, "y_is_mean"), V.validate_bool(obs_variance, "obs_variance")
        self.mu = V.validate_float(mu, "mu")
        self.sigma = V.validate_float_or_iterable_numerical(sigma, "sigma", positive=True)
        if self.gp_type in NYSTROEM_TYPES:
            self._fail(MSG_NO_NYSTROEM.format(gp_type=gp_type))

    def __call__(self, x=None, y=None):
        """``fit_predict(x, y)`` (``function_estimator.py:180-192``)."""
        return self.fit_predict(x=x, y=y)

    def __repr__(self):
        text = super().__repr__()
        return text[:-2] + f"\n    sigma={self.sigma},\n    y_is_mean={self.y_is_mean},\n)"

    def _repr_html_(self):
        unset = "Not Set"
        rows = {"Jitter": self.jitter, "Mean (μ)": self.mu or unset, "Length Scale (ls)": self.ls or unset,
                "Length-Scale Factor": self.ls_factor, "Noise Standard Deviation (σ)": self.sigma,
                "y_is_mean": self.y_is_mean, "Nearest Neighbor Distances": self.nn_distances}
        table = "".join(f"<tr><td>{k}</td><td>{object_html(v)}</td></tr>" for k, v in rows.items())
        status = "Available" if getattr(self, "conditional", None) else "Not Yet Computed"
        return (
            f"<div><h3>Function Estimator: {self.__class__.__name__}</h3>"
            f"<p>gp_type={self.gp_type}, n_landmarks={self.n_landmarks or 'Not Set'}, "
            f"covariance={object_html(self.cov_func or 'Not Set')}</p>"
            f"<table><tr><th>Parameter</th><th>Value</th></tr>{table}</table>"
            f"<p><strong>Predictor:</strong> {status}</p></div>"
        )

    def prepare_inference(self, x):
        """Fill n_landmarks, gp_type, (nn_distances,) ls, cov_func and landmarks
        (``function_estimator.py:295-316``)."""
        self.set_x(x)
        needs_distances = self.ls is None and self.cov_func is None     # only the length-scale heuristic reads them
        for name in ("n_landmarks", "gp_type", "nn_distances", "ls", "cov_func", "landmarks"):
            if name != "nn_distances" or needs_distances:
                self._prepare_attribute(name)

    def compute_conditional(self, x=None, y=None, obs_variance=None):
        """Condition the GP on ``y`` observed at ``x`` (``function_estimator.py:318-374``)."""
        given, x = x, (self.x if x is None else V.validate_array(x, "x"))
        if x is None:
            raise ValueError("Required argument x is missing and self.x has not been set.")
        if self.x is not None and self.x is not x and self._x_given is not given:
            logger.warning(MSG_OTHER_X)
        if y is None:
            raise ValueError("Required argument y is missing.")
        self.conditional = I.compute_conditional(
            x, self.landmarks, None, None, y, self.mu, self.cov_func, None, None, self.sigma, jitter=self.jitter,
            y_is_mean=self.y_is_mean, with_uncertainty=self.predictor_with_uncertainty,
            obs_variance=self.obs_variance if obs_variance is None else obs_variance,
        )
        return self.conditional

    def fit(self, x=None, y=None, obs_variance=None):
        """Prepare the covariance and landmarks, then condition on ``y`` (``function_estimator.py:376-420``)."""
        x, y = self.set_x(x), V.validate_array(y, "y")
        if y.shape[0] != x.shape[0]:
            raise ValueError(MSG_SAMPLES.format(n=x.shape[0], m=y.shape[0]))
        self.prepare_inference(x)
        self.compute_conditional(x, y, obs_variance=obs_variance)
        self.y = y
        return self

    @property
    def predict(self):
        """The fitted :class:`mellon_b200.Predictor` (``function_estimator.py:421-441``)."""
        return self.conditional

    def leverage(self, X=None):
        """Leverage with the fitted sigma, at the training points by default
        (``function_estimator.py:443-459``)."""
        return self.predict.leverage(self.x if X is None else X)

    def loo_residuals_squared(self, X=None, y=None):
        """Squared leave-one-out residuals ``r_i^2 / (1 - h_i)^2`` (``function_estimator.py:461-487``); without
        arguments, the ones kept from an ``obs_variance`` fit."""
        if X is None and y is None:
            if hasattr(self.predict, "_corrected_r2"):
                return self.predict._corrected_r2
            X, y = self.x, self.y
        else:
            X = self.x if X is None else X
            y = self.y if y is None else y
        return self.predict.loo_residuals_squared(X, y)

    def get_obs_variance(self, X=None):
        """Smoothed observation variance of the fitted predictor (``function_estimator.py:489-505``)."""
        return self.predict.obs_variance(self.x if X is None else X)

    def fit_predict(self, x=None, y=None, Xnew=None):
        """Fit, then return the conditional mean at ``Xnew`` (default: ``x``), one column per column of ``y``
        (``function_estimator.py:507-565``)."""
        x, y = self.set_x(x), V.validate_array(y, "y")
        Xnew = V.validate_array(Xnew, "Xnew", optional=True)
        if Xnew is not None and Xnew.ndim != x.ndim:
            raise ValueError(MSG_NDIM.format(a=x.ndim, b=Xnew.ndim))
        if Xnew is not None and x.ndim > 1 and Xnew.shape[1] != x.shape[1]:
            raise ValueError(MSG_FEATURES.format(a=x.shape[1], b=Xnew.shape[1]))
        self.fit(x, y)
        return self.predict(x if Xnew is None else Xnew)

    def multi_fit_predict(self, x=None, Y=None, Xnew=None):
        """Deprecated row-per-output form of :meth:`fit_predict` (``function_estimator.py:567-615``)."""
        logger.warning(MSG_DEPRECATED)
        x, Y = self.set_x(x), V.validate_array(Y, "Y")
        if Y.shape[0] != x.shape[0] and Y.shape[1] == x.shape[0]:
            logger.warning(MSG_TRANSPOSE)
            Y = Y.T
        return self.fit_predict(x, Y, Xnew).T
