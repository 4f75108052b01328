"""Estimator base class (``mellon/base_model.py``): constructor validation, the lazy
``_prepare_attribute`` pipeline, ``_compute_L`` with its rank checks and the optimiser switch.

Same constructor arguments, defaults, attributes and log / error text as the reference; the
attributes ``L`` and ``Lp`` are device handles (:class:`mellon_b200.backend.DeviceArray`) that
convert to NumPy on demand.
"""

from __future__ import annotations

import logging

from .cov import Matern52
from .inference import (DEFAULT_INIT_LEARN_RATE, DEFAULT_JIT, DEFAULT_N_ITER, DEFAULT_OPTIMIZER, compute_laplace_std,
                        minimize_adam, run_advi, minimize_lbfgsb)
from .parameter_validation import validate_cov_func, validate_cov_func_curry, validate_params
from .parameters import (DEFAULT_RANDOM_SEED, compute_cov_func, compute_gp_type, compute_L, compute_landmarks,
                         compute_Lp, compute_ls, compute_n_landmarks, compute_nn_distances, compute_rank)
from .util import DEFAULT_JITTER, GaussianProcessType, object_str, test_rank
from .validation import (validate_array, validate_bool, validate_float, validate_float_or_int,
                         validate_float_or_iterable_numerical, validate_nn_distances, validate_positive_float,
                         validate_positive_int, validate_string)

DEFAULT_COV_FUNC = Matern52
RANK_FRACTION_THRESHOLD = 0.8
SAMPLE_LANDMARK_RATIO = 10

logger = logging.getLogger("mellon")


class BaseEstimator:
    """Base class of the estimators (``base_model.py:56-482``)."""

    def __init__(self, cov_func_curry=DEFAULT_COV_FUNC, n_landmarks=None, rank=None, jitter=DEFAULT_JITTER,
                 optimizer=DEFAULT_OPTIMIZER, n_iter=DEFAULT_N_ITER, init_learn_rate=DEFAULT_INIT_LEARN_RATE,
                 landmarks=None, gp_type=None, nn_distances=None, d=None, mu=0, ls=None, ls_factor=1,
                 cov_func=None, Lp=None, L=None, initial_value=None, predictor_with_uncertainty=False,
                 jit=DEFAULT_JIT, check_rank=None, random_state=DEFAULT_RANDOM_SEED):
        self.cov_func_curry = validate_cov_func_curry(cov_func_curry, cov_func, "cov_func_curry")
        self.n_landmarks = validate_positive_int(n_landmarks, "n_landmarks", optional=True)
        self.random_state = validate_positive_int(random_state, "random_state", optional=True)
        self.rank = validate_float_or_int(rank, "rank", optional=True)
        self.jitter = validate_positive_float(jitter, "jitter")
        self.landmarks = validate_array(landmarks, "landmarks", optional=True)
        self.gp_type = GaussianProcessType.from_string(gp_type, optional=True)
        self.nn_distances = validate_array(nn_distances, "nn_distances", optional=True)
        self.nn_distances = validate_nn_distances(self.nn_distances, optional=True)
        self.mu = validate_float(mu, "mu", optional=True)
        self.ls = validate_positive_float(ls, "ls", optional=True)
        self.ls_factor = validate_positive_float(ls_factor, "ls_factor")
        self.cov_func = validate_cov_func(cov_func, "cov_func", optional=True)
        self.Lp = validate_array(Lp, "Lp", optional=True)
        self.L = validate_array(L, "L", optional=True)
        self.d = validate_float_or_iterable_numerical(d, "d", optional=True, positive=True)
        self.initial_value = validate_array(initial_value, "initial_value", optional=True)
        self.optimizer = validate_string(optimizer, "optimizer", choices={"adam", "advi", "L-BFGS-B"})
        self.n_iter = validate_positive_int(n_iter, "n_iter")
        self.init_learn_rate = validate_positive_float(init_learn_rate, "init_learn_rate")
        self.predictor_with_uncertainty = validate_bool(predictor_with_uncertainty, "predictor_with_uncertainty")
        self.jit = validate_bool(jit, "jit")
        self.check_rank = validate_bool(check_rank, "check_rank", optional=True)
        self.x = None
        self._x_given = None
        self.pre_transformation = None

    def __str__(self):
        return self.__repr__()

    def __repr__(self):
        fields = {
            "check_rank": self.check_rank,
            "cov_func": self.cov_func,
            "cov_func_curry": self.cov_func_curry,
            "d": object_str(self.d, ["cells"]),
            "gp_type": self.gp_type,
            "initial_value": object_str(self.initial_value, ["ranks"]),
            "jit": self.jit,
            "jitter": self.jitter,
            "landmarks": object_str(self.landmarks, ["landmarks", "dims"]),
            "L": object_str(self.L, ["cells", "ranks"]),
            "Lp": object_str(self.Lp, ["landmarks", "landmarks"]),
            "ls": self.ls,
            "ls_factor": self.ls_factor,
            "mu": self.mu,
            "n_landmarks": self.n_landmarks,
            "nn_distances": object_str(self.nn_distances, ["cells"]),
            "optimizer": self.optimizer,
            "predictor_with_uncertainty": self.predictor_with_uncertainty,
            "random_state": self.random_state,
            "rank": self.rank,
        }
        body = "".join(f"\n    {key}={value}," for key, value in fields.items())
        return f"{self.__class__.__name__}({body}\n)"

    def __call__(self, x=None):
        """Fit the model and predict on the training data (``base_model.py:165-174``)."""
        return self.fit_predict(x=x)

    @staticmethod
    def _fail(message):
        error = ValueError(message)
        logger.error(error)
        raise error

    def set_x(self, x):
        """Validate and store the training instances (``base_model.py:176-213``).  Passing a
        different object than the one already set raises — identity, not equality."""
        if self.x is not None and x is not None and self.x is not x and getattr(self, "_x_given", None) is not x:
            self._fail("self.x has been set already, but is not equal to the argument x.")
        if self.x is None and x is None:
            self._fail("Required argument x is missing and self.x has not been set.")
        if x is None:
            x = self.x
        elif x is not self.x:
            # validation turns ndarray subclasses / foreign array types into a NEW float64 array; remember the object
            # the caller passed, so that passing it again is still "the same x" as in the reference (jnp.asarray keeps it)
            self._x_given = x
        self.x = validate_array(x, "x")
        return self.x

    # -- lazy attribute pipeline ------------------------------------------------------------
    def _compute_n_landmarks(self):
        return compute_n_landmarks(self.gp_type, self.x.shape[0], self.landmarks)

    def _seed(self):
        return self.random_state if self.random_state is not None else DEFAULT_RANDOM_SEED

    def _compute_landmarks(self):
        n_samples, n_landmarks = self.x.shape[0], self.n_landmarks
        if n_samples > 100 * n_landmarks and n_samples > 1e6:
            logger.info(
                f"Large number of {n_samples:,} cells and small number of {n_landmarks:,} landmarks. "
                "Consider computing k-means on a subset of cells and passing the results as "
                "'landmarks' to speed up the process."
            )
        return compute_landmarks(self.x, self.gp_type, n_landmarks=n_landmarks, random_state=self._seed())

    def _compute_rank(self):
        return compute_rank(self.gp_type)

    def _compute_gp_type(self):
        return compute_gp_type(self.n_landmarks, self.rank, self.x.shape[0])

    def _compute_nn_distances(self):
        logger.info("Computing nearest neighbor distances.")
        return validate_nn_distances(compute_nn_distances(self.x, seed=self._seed()))

    def _compute_ls(self):
        return compute_ls(self.nn_distances) * self.ls_factor

    def _compute_cov_func(self):
        cov_func = compute_cov_func(self.cov_func_curry, self.ls)
        logger.info("Using covariance function %s.", str(cov_func))
        return cov_func

    def _compute_Lp(self):
        return compute_Lp(self.x, self.cov_func, self.gp_type, self.landmarks, sigma=0, jitter=self.jitter)

    def _compute_L(self):
        """``compute_L`` plus the two sanity checks of ``base_model.py:299-358``."""
        x, landmarks, gp_type, rank = self.x, self.landmarks, self.gp_type, self.rank
        L = compute_L(x, self.cov_func, gp_type, landmarks=landmarks, Lp=self.Lp, rank=rank, sigma=0,
                      jitter=self.jitter)
        new_rank = L.shape[1]
        n_samples = x.shape[0]
        n_landmarks = n_samples if landmarks is None else landmarks.shape[0]
        nystroem = gp_type in (GaussianProcessType.SPARSE_NYSTROEM, GaussianProcessType.FULL_NYSTROEM)
        if nystroem and new_rank > (rank * RANK_FRACTION_THRESHOLD * n_landmarks):
            logger.warning(
                f"Shallow rank reduction from {n_landmarks:,} to {new_rank:,} indicates "
                "underrepresentation by landmarks. Consider increasing n_landmarks!"
            )
        check_rank = self.check_rank
        if (
            check_rank is None
            and gp_type == GaussianProcessType.SPARSE_CHOLESKY
            and SAMPLE_LANDMARK_RATIO * n_landmarks < n_samples
        ) or (check_rank is not None and check_rank):
            logger.info(
                f"Estimating approximation accuracy since {n_samples:,} samples are more than "
                f"{SAMPLE_LANDMARK_RATIO} x {n_landmarks:,} landmarks."
            )
            test_rank(L, threshold=RANK_FRACTION_THRESHOLD)
        logger.info(f"Using rank {new_rank:,} covariance representation.")
        return L

    def validate_parameter(self):
        """No contradictions between rank, gp_type and landmarks (``base_model.py:360-369``)."""
        validate_params(self.rank, self.gp_type, self.x.shape[0], self.n_landmarks, self.landmarks)

    def _run_inference(self):
        """Optimiser switch + Laplace hook (``base_model.py:371-431``)."""
        function, initial_value, optimizer = self.loss_func, self.initial_value, self.optimizer
        logger.info("Running inference using %s.", optimizer)
        if optimizer == "adam":
            results = minimize_adam(function, initial_value, n_iter=self.n_iter,
                                    init_learn_rate=self.init_learn_rate, jit=self.jit)
            self.pre_transformation = results.pre_transformation
            self.pre_transformation_std = None
            self.opt_state = results.opt_state
            self.losses = results.losses
        elif optimizer == "advi":
            results = run_advi(function, initial_value, n_iter=self.n_iter, init_learn_rate=self.init_learn_rate,
                               jit=self.jit)
            self.pre_transformation = results.pre_transformation
            self.pre_transformation_std = results.pre_transformation_std
            self.losses = results.losses
        elif optimizer == "L-BFGS-B":
            results = minimize_lbfgsb(function, initial_value, jit=self.jit)
            self.pre_transformation = results.pre_transformation
            self.pre_transformation_std = None
            self.opt_state = results.opt_state
            self.losses = [results.loss]
        else:
            self._fail(
                f"Unknown optimizer {optimizer}. You can use .loss_func and "
                ".initial_value as loss function and initial state for an "
                "external optimization. Write optimal state to "
                ".pre_transformation to enable prediction with .predict()."
            )
        if optimizer != "advi" and self.predictor_with_uncertainty and self.pre_transformation_std is None:
            logger.info("Computing Laplace approximation for posterior uncertainty.")
            self.pre_transformation_std = compute_laplace_std(function, self.pre_transformation, jit=self.jit)

    def _prepare_attribute(self, attribute):
        """Fill ``self.<attribute>`` from ``_compute_<attribute>`` unless it is already set
        (``base_model.py:433-446``)."""
        if getattr(self, attribute) is not None:
            return
        from .backend import get_backend

        with get_backend().range("prepare " + attribute):   # NVTX: one range per stage of prepare_inference
            setattr(self, attribute, getattr(self, "_compute_" + attribute)())

    def prepare_inference(self, x):  # pragma: no cover - interface
        ...

    def fit(self):  # pragma: no cover - interface
        ...

    @property
    def predict(self):  # pragma: no cover - interface
        ...

    def fit_predict(self, x):  # pragma: no cover - interface
        ...
