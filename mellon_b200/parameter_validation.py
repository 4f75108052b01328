"""Consistency checks between gp_type, rank and landmarks (``mellon/parameter_validation.py``)."""

from __future__ import annotations

import logging

import numpy as np

from .base_cov import Covariance
from .util import GaussianProcessType
from .validation import validate_float_or_int, validate_positive_int

logger = logging.getLogger("mellon")

_NYSTROEM = (GaussianProcessType.FULL_NYSTROEM, GaussianProcessType.SPARSE_NYSTROEM)


def _fail(message, level=logging.ERROR):
    logger.log(level, message)
    raise ValueError(message)


def validate_landmark_params(n_landmarks, landmarks):
    """``parameter_validation.py:14-40``"""
    if landmarks is not None and n_landmarks != landmarks.shape[0]:
        _fail(
            f"There are {landmarks.shape[0]:,} landmarks specified but n_landmarks={n_landmarks:,}. "
            "Please omit specifying n_landmarks if landmarks are given."
        )


def validate_rank_params(gp_type, n_samples, rank, n_landmarks):
    """``parameter_validation.py:43-100`` — a "full" rank contradicts a Nystroem type and a
    reduced rank contradicts a non-Nystroem type."""
    limit = n_landmarks if gp_type in (GaussianProcessType.SPARSE_CHOLESKY,
                                       GaussianProcessType.SPARSE_NYSTROEM) else n_samples
    int_full = type(rank) is int and gp_type in (
        GaussianProcessType.SPARSE_CHOLESKY, GaussianProcessType.SPARSE_NYSTROEM,
        GaussianProcessType.FULL, GaussianProcessType.FULL_NYSTROEM) and rank >= limit
    full_rank = int_full or (type(rank) is float and rank >= 1.0) or rank == 0
    if full_rank:
        if gp_type == GaussianProcessType.FULL_NYSTROEM:
            _fail(
                f"Gaussian Process type {gp_type} requires fractional 0 < rank < 1 or integer "
                f"0 < rank < {n_samples:,} (number of cells) but the actual rank is {rank}."
            )
        if gp_type == GaussianProcessType.SPARSE_NYSTROEM:
            _fail(
                f"Gaussian Process type {gp_type} requires fractional 0 < rank < 1 or integer "
                f"0 < rank < {n_landmarks:,} (number of landmakrs) but the actual rank is {rank}."
            )
    elif gp_type not in _NYSTROEM:
        _fail(
            f"Given rank {rank} indicates Nyström rank reduction. "
            f"But the Gaussian Process type is set to {gp_type}."
        )


def validate_gp_type(gp_type, n_samples, n_landmarks):
    """``parameter_validation.py:103-146``"""
    if gp_type in (GaussianProcessType.FULL, GaussianProcessType.FULL_NYSTROEM):
        if n_landmarks != 0 and n_landmarks < n_samples:
            _fail(
                f"Gaussian Process type {gp_type} but n_landmarks={n_landmarks:,} is smaller "
                f"than the number of cells {n_samples:,}. Omit n_landmarks or set it to 0 to use "
                "a non-sparse Gaussian Process or omit gp_type to use a sparse one."
            )
    elif gp_type in (GaussianProcessType.SPARSE_CHOLESKY, GaussianProcessType.SPARSE_NYSTROEM):
        if n_landmarks == 0:
            _fail(
                f"Gaussian Process type {gp_type} but n_landmarks=0. Set n_landmarks "
                f"to a number smaller than the number of cells {n_samples:,} to use a"
                "sparse Gaussuian Process or omit gp_type to use a non-sparse one."
            )
        elif n_landmarks >= n_samples:
            _fail(
                f"Gaussian Process type {gp_type} but n_landmarks={n_landmarks:,} is larger or "
                f"equal the number of cells {n_samples:,}. Reduce the number of landmarks to use a"
                "sparse Gaussuian Process or omit gp_type to use a non-sparse one.",
                level=logging.WARNING,
            )


def validate_params(rank, gp_type, n_samples, n_landmarks, landmarks):
    """``parameter_validation.py:149-192``"""
    n_landmarks = validate_positive_int(n_landmarks, "n_landmarks")
    rank = validate_float_or_int(rank, "rank")
    if not isinstance(gp_type, GaussianProcessType):
        _fail(f"gp_type needs to be a mellon.util.GaussianProcessType but is a {type(gp_type)} instead.")
    validate_landmark_params(n_landmarks, landmarks)
    if n_landmarks > n_samples and gp_type != GaussianProcessType.FIXED:
        logger.warning(f"n_landmarks={n_landmarks:,} is larger than the number of cells {n_samples:,}.")
    validate_gp_type(gp_type, n_samples, n_landmarks)
    validate_rank_params(gp_type, n_samples, rank, n_landmarks)


def validate_cov_func_curry(cov_func_curry, cov_func, param_name):
    """``parameter_validation.py:195-226``"""
    if cov_func_curry is None and cov_func is None:
        raise ValueError("At least one of 'cov_func_curry' and 'cov_func' must not be None")
    if cov_func_curry is not None:
        if not isinstance(cov_func_curry, type) or not issubclass(cov_func_curry, Covariance):
            raise ValueError(f"'{param_name}' must be a subclass of mellon.Covariance")
    return cov_func_curry


def validate_cov_func(cov_func, param_name, optional=False):
    """``parameter_validation.py:229-255``"""
    if cov_func is None and optional:
        return None
    if not isinstance(cov_func, Covariance):
        raise ValueError(f"'{param_name}' must be an instance of a subclass of mellon.Covariance")
    return cov_func


def validate_normalize_parameter(normalize, unique_times):
    """``parameter_validation.py:258-279``"""
    if isinstance(normalize, dict):
        missing = [t for t in unique_times if t.item() not in normalize]
        if missing:
            raise ValueError(f"Missing time point(s) in normalization dictionary: {missing}")
    elif isinstance(normalize, (list, np.ndarray)) and len(normalize) != len(unique_times):
        raise ValueError("Length of the normalize list or array must match the number of unique time points.")
