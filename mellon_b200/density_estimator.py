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

The steps shared with the time-sensitive estimator are in :mod:`mellon_b200.density_pipeline`.
"""

from __future__ import annotations

from .base_model import DEFAULT_COV_FUNC
from .density_pipeline import DEFAULT_D_METHOD, DensityPipeline
from .inference import DEFAULT_INIT_LEARN_RATE, DEFAULT_JIT, DEFAULT_N_ITER, DEFAULT_OPTIMIZER, compute_conditional
from .parameters import DEFAULT_RANDOM_SEED
from .util import DEFAULT_JITTER
from .validation import validate_array


class DensityEstimator(DensityPipeline):
    """Non-parametric density estimator: a GP prior on the log density and the
    nearest-neighbour-distance likelihood (``density_estimator.py:38-581``)."""

    # density_estimator.py:432-444: the neighbour distances come before d
    PIPELINE = ("nn_distances", "d", "mu", "ls", "cov_func", "landmarks", "Lp", "L", "initial_value", "transform",
                "loss_func")

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
        self._init_density(d, d_method)

    def _repr_html_(self):
        unset = "Not Set"
        return self._html({
            "Jitter": self.jitter, "Mean (μ)": self.mu or unset, "Length Scale (ls)": self.ls or unset,
            "Length-Scale Factor": self.ls_factor, "Dimensionality (d)": unset if self.d is None else self.d,
            "Nearest Neighbor Distances": self.nn_distances, "Landmarks": self.landmarks, "L": self.L, "Lp": self.Lp,
        })

    def _set_log_density_func(self):
        """Build the predictor (``density_estimator.py:370-402``)."""
        self._build_predictor(compute_conditional, n_obs=self.x.shape[0])

    def prepare_inference(self, x):
        """Compute every missing parameter; returns ``(loss_func, initial_value)``
        (``density_estimator.py:404-444``)."""
        return self._prepare_pipeline(self._claim_x(x))

    def fit(self, x=None, build_predict=True):
        """Fit the model from end to end (``:494-516``)."""
        self.prepare_inference(x)
        self.run_inference()
        self.process_inference(build_predict=build_predict)
        return self

    def fit_predict(self, x=None, build_predict=False):
        """Fit and return the log density at the training points (``:542-581``)."""
        x = self._claim_x(x, validate=lambda a: validate_array(a, "x"), validate_first=False)
        self.fit(x, build_predict=build_predict)
        return self.log_density_x
