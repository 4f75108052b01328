"""Predictor shells (``mellon/base_predictor.py``): input validation, normalisation,
serialisation.  The arithmetic (``_mean`` etc.) lives in :mod:`mellon_b200.conditional` and
runs on the GPU.
"""

from __future__ import annotations

import bz2
import gzip
import json
import logging
import sys
from abc import ABC, abstractmethod
from datetime import datetime
from importlib import import_module
from typing import List, Set, Union

import numpy as np
from packaging import version

from .base_cov import Covariance, wire_module_name
from .util import deserialize, ensure_2d, make_multi_time_argument, make_serializable, object_html, object_str
from .validation import validate_array, validate_bool, validate_time_x

logger = logging.getLogger("mellon")


def _feature_error(expected, got):
    return ValueError(
        f"The predictor was trained on data with {expected} features. "
        f"However, the provided input data has {got} features. "
        "Please ensure that the input data has the same number of features as the training data."
    )



def _central_gradient(f, x, cols, h_rel=1e-5):
    """d f / d x[:, j] for j in ``cols`` by central differences; ``f`` maps (n, d) -> (n,).  Each column costs two
    evaluations of the (device) predictor mean."""
    x = np.asarray(x, dtype=float)
    out = None
    for a, j in enumerate(cols):
        h = h_rel * max(1.0, float(np.max(np.abs(x[:, j])))) if x.shape[0] else h_rel
        e = np.zeros(x.shape[1])
        e[j] = h
        diff = (np.asarray(f(x + e)) - np.asarray(f(x - e))) / (2.0 * h)
        if out is None:   # (n, k) for a scalar mean, (n, p, k) for a multi-output one (FunctionEstimator with y of shape (n, p))
            out = np.empty(diff.shape + (len(cols),))
        out[..., a] = diff
    return out if out is not None else np.empty((x.shape[0], 0))


def _central_hessian(f, x, cols, h_rel=1e-4):
    """Second derivatives over the columns ``cols`` by central differences: (n, len(cols), len(cols))."""
    x = np.asarray(x, dtype=float)
    k = len(cols)
    hs = [h_rel * max(1.0, float(np.max(np.abs(x[:, j])))) if x.shape[0] else h_rel for j in cols]
    f0 = np.asarray(f(x))
    if f0.ndim != 1:
        raise NotImplementedError("hessian / hessian_log_determinant of a multi-output predictor (a FunctionEstimator fitted "
                                  "with y of shape (n, p)) are not provided; take them per output column.")
    out = np.empty((x.shape[0], k, k))
    for a, j in enumerate(cols):
        ej = np.zeros(x.shape[1])
        ej[j] = hs[a]
        out[:, a, a] = (np.asarray(f(x + ej)) - 2.0 * f0 + np.asarray(f(x - ej))) / hs[a] ** 2
        for b in range(a + 1, k):
            ek = np.zeros(x.shape[1])
            ek[cols[b]] = hs[b]
            v = (np.asarray(f(x + ej + ek)) - np.asarray(f(x + ej - ek)) - np.asarray(f(x - ej + ek))
                 + np.asarray(f(x - ej - ek))) / (4.0 * hs[a] * hs[b])
            out[:, a, b] = out[:, b, a] = v
    return out


class Predictor(ABC):
    """Callable posterior-mean predictor (``base_predictor.py:41-734``)."""

    n_input_features: int
    n_obs: int
    d: int = None
    d_method: str = None
    _state_variables: Union[Set, List]

    @abstractmethod
    def __init__(self):
        """Subclasses set cov_func, weights, mu, n_input_features, n_obs, _state_variables."""

    def __str__(self):
        return self.__repr__()

    def __repr__(self):
        n_obs = "None" if self.n_obs is None else f"{self.n_obs:,}"
        rows = "\n".join(str(k) + ": " + object_str(v) for k, v in self._data_dict().items())
        return (
            f'A predictor of class "{self.__class__.__name__}" with covariance function '
            f'"{self.cov_func!r}" trained on {n_obs} observations '
            f"with {self.n_input_features:,} features and data:\n{rows}"
        )

    def _repr_html_(self):
        n_obs = "None" if self.n_obs is None else f"{self.n_obs:,}"
        rows = "".join(f"<tr><td>{k}</td><td>{object_html(v)}</td></tr>" for k, v in self._data_dict().items())
        return (
            f"<div><h3>Predictor: {self.__class__.__name__}</h3><p>Covariance: {self.cov_func!r}; "
            f"{n_obs} observations; {self.n_input_features:,} features</p>"
            f"<table><tr><th>Attribute</th><th>Value</th></tr>{rows}</table></div>"
        )

    # -- evaluation -----------------------------------------------------------------------------
    @abstractmethod
    def _mean(self, *args, **kwargs):
        """Posterior mean at validated inputs."""

    def _check_x(self, x):
        x = ensure_2d(validate_array(x, "x"))
        if x.shape[1] != self.n_input_features:
            raise _feature_error(self.n_input_features, x.shape[1])
        return x

    def _normalisation_notes(self, what="samples/cells"):
        if self.n_obs is None or self.n_obs == 0:
            message = (
                f"Cannot normalize without n_obs. Please set self.n_obs to the number "
                f"of {what} trained on to enable normalization."
            )
            logger.error(message)
            raise ValueError(message)
        if self.d_method == "manual":
            logger.info(
                f"Using normalization with manually set d={self.d}. "
                "Note: Normalization is most effective when d approximates the intrinsic dimensionality of the data."
            )
        elif self.d_method == "embedding" or (
            self.d_method is None and isinstance(self.d, (int, float)) and float(self.d).is_integer()
        ):
            logger.warning(
                f"The normalization is only effective if d approximates the intrinsic dimensionality. "
                f"Current values: d_method={self.d_method}, d={self.d}. "
                f'Consider using d_method="fractal" for more accurate results.'
            )

    def mean(self, x, normalize=False):
        """Posterior mean at ``x`` (``base_predictor.py:180-255``); ``normalize`` subtracts
        ``log(n_obs)``."""
        x = self._check_x(x)
        normalize = validate_bool(normalize, "normalize")
        if normalize:
            self._normalisation_notes()
            return self._mean(x) - np.log(self.n_obs)
        return self._mean(x)

    __call__ = mean

    # -- regression diagnostics (FunctionEstimator; base_predictor.py:259-355) -----------------------------
    def _leverage(self, Xnew, sigma):
        raise NotImplementedError(f"{self.__class__.__name__} does not define a leverage.")

    def leverage(self, x):
        """Diagonal of the GP hat matrix ``H = K (K + sigma^2 I)^-1`` (``base_predictor.py:263-288``):
        shape ``(n,)``, or ``(n, p)`` with a per-feature sigma."""
        return self._leverage(self._check_x(x), self.sigma)

    def loo_residuals_squared(self, x, y):
        """Squared leave-one-out residuals through the HC3 shortcut ``r_i^2 / (1 - h_i)^2``
        (``base_predictor.py:290-325``)."""
        x = validate_array(x, "x")
        y = validate_array(y, "y")
        x = self._check_x(x)
        residual = y - self._mean(x)
        h = self._leverage(x, self.sigma)
        if residual.ndim > h.ndim:
            h = h[..., None]
        return residual ** 2 / (1 - h) ** 2

    def _obs_variance(self, Xnew):
        raise NotImplementedError(f"{self.__class__.__name__} does not define an observation variance.")

    def obs_variance(self, x):
        """GP-smoothed corrected squared residuals: an input-dependent estimate of the observation noise
        variance (``base_predictor.py:330-355``)."""
        return self._obs_variance(self._check_x(x))

    @abstractmethod
    def _covariance(self, *args, **kwargs):
        """Posterior covariance of the GP at validated inputs."""

    @abstractmethod
    def _mean_covariance(self, *args, **kwargs):
        """Covariance of the mean induced by parameter uncertainty."""

    def _has_per_feature_sigma(self):
        return getattr(self, "per_feature_sigma", False)

    def covariance(self, x, diag=True, noise_free=False):
        """Posterior (co)variance (``base_predictor.py:361-418``).  A predictor fitted with a per-feature sigma
        only has the noise-free covariance, which the caller must acknowledge with ``noise_free=True``."""
        if self._has_per_feature_sigma() and not noise_free:
            raise ValueError(
                "This predictor was fitted with per-feature sigma, so the "
                "covariance is noise-free (sigma=0) and does not include "
                "observation noise. Pass noise_free=True to acknowledge this "
                "and obtain the noise-free covariance, then account for "
                "observation noise separately (e.g., via obs_variance)."
            )
        return self._covariance(self._check_x(x), diag=diag)

    def mean_covariance(self, x, diag=True):
        return self._mean_covariance(self._check_x(x), diag=diag)

    def uncertainty(self, x, diag=True):
        x = self._check_x(x)
        return self._covariance(x, diag=diag) + self._mean_covariance(x, diag=diag)

    # -- derivatives of the mean (base_predictor.py:490-539).  The reference differentiates with JAX; here the
    # device-evaluated mean is differenced centrally (relative error ~1e-9 for the gradient, ~1e-6 for the
    # Hessian): an API-compatible host utility, not part of the accelerated path.
    def gradient(self, x, jit=True):
        x = ensure_2d(self._check_x(x))
        return _central_gradient(self._mean, x, list(range(x.shape[1])))

    def hessian(self, x, jit=True):
        x = ensure_2d(self._check_x(x))
        return _central_hessian(self._mean, x, list(range(x.shape[1])))

    def hessian_log_determinant(self, x, jit=True):
        sign, logdet = np.linalg.slogdet(self.hessian(x))
        return sign, logdet

    def _data_dict(self):
        return {key: getattr(self, key) for key in self._state_variables}

    # -- serialisation (wire format of base_predictor.py:541-734) ------------------------------
    def __getstate__(self):
        module_name = self.__class__.__module__
        meta = import_module(module_name.split(".")[0])
        data = self._data_dict()
        data.update(
            {
                "n_input_features": self.n_input_features,
                "n_obs": self.n_obs,
                "d": self.d,
                "d_method": self.d_method,
                "_state_variables": self._state_variables,
            }
        )
        return {
            "data": {k: make_serializable(v) for k, v in data.items()},
            "cov_func": self.cov_func.__getstate__(),
            "metadata": {
                "classname": self.__class__.__name__,
                "module_name": wire_module_name(module_name),
                "module_version": getattr(meta, "__version__", "NA"),
                "serialization_date": datetime.now().isoformat(),
                "python_version": sys.version,
            },
        }

    def __setstate__(self, state):
        for name, value in state["data"].items():
            setattr(self, name, deserialize(value))
        self.cov_func = Covariance.from_dict(state["cov_func"])

    def copy(self):
        new = self.__class__.__new__(self.__class__)
        new.__setstate__(self.__getstate__())
        return new

    def to_dict(self):
        return self.__getstate__()

    def to_json(self, filename=None, compress=None):
        json_str = json.dumps(self.to_dict())
        if filename is None:
            return json_str
        if compress == "gzip":
            if isinstance(filename, str) and not filename.endswith(".gz"):
                filename += ".gz"
            with gzip.open(filename, "wt") as f:
                f.write(json_str)
        elif compress == "bz2":
            if isinstance(filename, str) and not filename.endswith(".bz2"):
                filename += ".bz2"
            with bz2.open(filename, "wt") as f:
                f.write(json_str)
        elif compress is None:
            with open(filename, "w") as f:
                f.write(json_str)
        else:
            msg = f'Unknown compression format {compress}.\nAvailabe formats are "gzip", "bz2" and None.'
            logger.error(msg)
            raise ValueError(msg)
        logger.info(f"Written predictor to {filename}.")

    @classmethod
    def from_json(cls, filepath, compress=None):
        name = str(filepath)
        if compress == "gzip" or name.endswith(".gz"):
            opener = gzip.open
        elif compress == "bz2" or name.endswith(".bz2"):
            opener = bz2.open
        else:
            opener = open
        with opener(filepath, "rt") as f:
            return cls.from_json_str(f.read())

    @classmethod
    def from_json_str(cls, json_str):
        return cls.from_dict(json.loads(json_str))

    @classmethod
    def from_dict(cls, data_dict):
        meta = data_dict["metadata"]
        clsname, module_name = meta["classname"], meta["module_name"]
        try:
            old = version.parse(str(meta["module_version"])) < version.parse("1.4.0")
        except version.InvalidVersion:
            old = False
        if old and module_name.startswith("mellon."):
            logger.warning(
                f"Loading a predictor written by mellon {meta['module_version']} < 1.4.0. "
                "Please set predictor.n_obs to enable normalization."
            )
            clsname = clsname.replace("ConditionalMean", "Conditional")
            data = data_dict["data"]
            data["n_obs"] = data.get("n_obs", None)
            data["_state_variables"] = data.get("_state_variables", set(data.keys()) - {"n_input_features"})
        # predictors written by the reference name its module ("mellon.conditional"): same classes here.  The
        # class-name shortcut only applies to the reference's and this package's own modules: a user class that merely
        # shares a name with a stock one is imported from the module it names.
        from . import conditional

        ours = module_name.split(".")[0] in ("mellon", "mellon_b200")
        if ours and clsname.startswith("Exp"):
            raise NotImplementedError(
                f"{clsname} (the reference's exp-log predictor family, mellon/conditional.py) is outside the path "
                "mellon_b200 implements; load this file with mellon.")
        Sub = getattr(conditional, clsname, None) if ours else None
        if Sub is None:
            Sub = getattr(import_module(module_name), clsname)
        instance = Sub.__new__(Sub)
        instance.__setstate__(data_dict)
        return instance


class PredictorTime(Predictor):
    """Predictor whose last input column is time (``base_predictor.py:852-1194``)."""

    def _check_time_x(self, Xnew, time):
        return validate_time_x(Xnew, time, n_features=self.n_input_features, cast_scalar=True)

    @make_multi_time_argument
    def mean(self, Xnew, time=None, normalize=False):
        Xnew = self._check_time_x(Xnew, time)
        normalize = validate_bool(normalize, "normalize")
        if normalize:
            self._normalisation_notes("samples/cells (per time point)")
            return self._mean(Xnew) - np.log(self.n_obs)
        return self._mean(Xnew)

    __call__ = mean

    @make_multi_time_argument
    def covariance(self, Xnew, time=None, diag=True):
        return self._covariance(self._check_time_x(Xnew, time), diag=diag)

    @make_multi_time_argument
    def mean_covariance(self, Xnew, time=None, diag=True):
        return self._mean_covariance(self._check_time_x(Xnew, time), diag=diag)

    @make_multi_time_argument
    def uncertainty(self, Xnew, time=None, diag=True):
        Xnew = self._check_time_x(Xnew, time)
        return self._covariance(Xnew, diag=diag) + self._mean_covariance(Xnew, diag=diag)

    # -- derivatives (base_predictor.py:1052-1194): with respect to the state columns at fixed time, and with
    # respect to time; central differences of the device-evaluated mean, as in Predictor.gradient
    @make_multi_time_argument
    def time_derivative(self, x, time=None, jit=True):
        Xnew = np.asarray(self._check_time_x(x, time), dtype=float)
        return _central_gradient(self._mean, Xnew, [Xnew.shape[1] - 1])[:, 0]

    @make_multi_time_argument
    def gradient(self, x, time=None, jit=True):
        Xnew = np.asarray(self._check_time_x(x, time), dtype=float)
        return _central_gradient(self._mean, Xnew, list(range(Xnew.shape[1] - 1)))

    @make_multi_time_argument
    def hessian(self, x, time=None, jit=True):
        Xnew = np.asarray(self._check_time_x(x, time), dtype=float)
        return _central_hessian(self._mean, Xnew, list(range(Xnew.shape[1] - 1)))

    @make_multi_time_argument
    def hessian_log_determinant(self, x, time=None, jit=True):
        sign, logdet = np.linalg.slogdet(self.hessian(x, time))
        return sign, logdet
