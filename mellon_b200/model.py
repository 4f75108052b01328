"""Estimator classes (``mellon/model.py``)."""

from .density_estimator import DensityEstimator
from .function_estimator import FunctionEstimator
from .time_sensitive_density_estimator import TimeSensitiveDensityEstimator

__all__ = ["DensityEstimator", "FunctionEstimator", "TimeSensitiveDensityEstimator"]
