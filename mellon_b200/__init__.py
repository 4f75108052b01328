"""mellon_b200 — B200-native drop-in for the sparse-GP density path of settylab/Mellon.

``import mellon_b200 as mellon`` gives the surface of ``mellon/__init__.py`` for that path:
``DensityEstimator``, ``TimeSensitiveDensityEstimator``, ``FunctionEstimator``, ``Predictor``, ``Covariance`` and the
sub-modules ``cov``, ``util``, ``parameters``, ``inference``, ``conditional``,
``decomposition``, ``validation``.  Importing the package needs neither a GPU nor the shared
library; the first numeric call does, and fails loudly without them (no CPU fallback).
"""

import logging.config
import sys

from .version import __version__

LOGGING_CONFIG = {
    "version": 1,
    "disable_existing_loggers": False,
    "formatters": {"standard": {"format": "[%(asctime)s] [%(levelname)-8s] %(message)s"}},
    "handlers": {
        "console": {
            "level": "DEBUG",
            "class": "logging.StreamHandler",
            "formatter": "standard",
            "stream": sys.stdout,
        },
    },
    "loggers": {"mellon": {"handlers": ["console"], "level": "INFO", "propagate": False}},
}


def setup_logging(config=None):
    """Configure the ``"mellon"`` logger (``mellon/__init__.py:63-105``)."""
    logging.config.dictConfig(LOGGING_CONFIG if config is None else config)
    return logging.getLogger("mellon")


def setup_jax(enable_x64=True, platform="cpu"):
    """Kept for source compatibility (``mellon/__init__.py:44-55``): there is no JAX here.
    Arithmetic is always float64 on the GPU; anything else is refused."""
    if not enable_x64:
        raise ValueError("mellon_b200 computes in float64 only.")


logger = setup_logging()

from . import conditional, cov, decomposition, inference, parameters, util, validation  # noqa: E402
from .backend import get_backend, set_backend  # noqa: E402
from .base_cov import Covariance  # noqa: E402
from .base_predictor import Predictor  # noqa: E402
from .model import DensityEstimator, FunctionEstimator, TimeSensitiveDensityEstimator  # noqa: E402

__all__ = [
    "DensityEstimator",
    "FunctionEstimator",
    "TimeSensitiveDensityEstimator",
    "Predictor",
    "Covariance",
    "util",
    "cov",
    "model",
    "parameters",
    "inference",
    "conditional",
    "decomposition",
    "validation",
    "__version__",
    "setup_jax",
    "setup_logging",
    "get_backend",
    "set_backend",
]
