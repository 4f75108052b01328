"""Argument validation with the reference's error contract (``mellon/validation.py``).

Same function names, same exception types and the same message text the reference's tests
assert on.  Arrays come back as NumPy float64; an input that already is a float64 ndarray (or
a device handle) is returned unchanged, which the estimators' ``self.x is x`` identity rule
(``mellon/base_model.py:200-204``) relies on.
"""

from __future__ import annotations

import logging
from collections.abc import Iterable

import numpy as np

logger = logging.getLogger("mellon")


def _unwrap_scalar(value, only_0d=False):
    """NumPy / device scalars -> Python scalars (so isinstance checks behave)."""
    item = getattr(value, "item", None)
    if callable(item) and (not only_0d or getattr(value, "ndim", None) == 0):
        try:
            return item()
        except (ValueError, TypeError):
            return value
    return value


def validate_array(iterable, name, optional=False, ndim=None):
    """``validation.py:302-361`` — to float64 array; TypeError for None / non-iterables."""
    from .backend import DeviceArray

    if iterable is None:
        if optional:
            return None
        raise TypeError(f"'{name}' can't be None.")
    if isinstance(iterable, DeviceArray):
        array = iterable
    elif hasattr(iterable, "todense"):
        array = np.asarray(iterable.todense(), dtype=float)
    elif isinstance(iterable, Iterable):
        array = np.asarray(iterable, dtype=float)
    else:
        raise TypeError(f"'{name}' should be iterable or sparse, got {type(iterable)} instead.")
    if ndim is not None:
        allowed = (ndim,) if isinstance(ndim, int) else tuple(ndim)
        if array.ndim not in allowed:
            raise ValueError(
                f"'{name}' must be a {allowed}-dimensional array, got {array.ndim}-dimensional array instead."
            )
    return array


def validate_time_x(x, times=None, n_features=None, cast_scalar=False):
    """``validation.py:23-101`` — append ``times`` as the last column of ``x``."""
    x = validate_array(x, "x", ndim=2)
    if cast_scalar and times is not None and (np.isscalar(times) or all(s == 1 for s in np.shape(times))):
        times = np.full(x.shape[0], np.asarray(times, dtype=float).reshape(-1)[0])
    times = validate_array(times, "times", optional=True, ndim=(1, 2))
    if times is not None:
        if times.ndim == 1:
            times = times.reshape(-1, 1)
        elif times.ndim != 2 or times.shape[1] != 1:
            raise ValueError("'times' must be a 1D array or a 2D array with 1 column.")
        if x.shape[0] != times.shape[0]:
            raise ValueError(
                "'x' and 'times' must have the same number of samples. "
                f"Got {x.shape[0]} for 'x' and {times.shape[0]} for 'times'."
            )
        x = np.concatenate((x, times), axis=1)
    if n_features is not None:
        if x.shape[1] == n_features - 1 and times is None:
            raise ValueError(
                f"Expected {n_features} features including 'times' in 'x' but "
                f"only found {x.shape[1]} features and 'times' is not provided."
            )
        if x.shape[1] != n_features:
            raise ValueError(f"Wrong number of features in 'x'. Expected {n_features} but got {x.shape[1]}.")
    return x


def validate_float_or_int(value, param_name, optional=False):
    """``validation.py:104-147``"""
    if value is None and optional:
        return None
    value = _unwrap_scalar(value)
    if not isinstance(value, (float, int)):
        try:
            value = float(value)
        except TypeError:
            raise ValueError(
                f"'{param_name}' should be a positive integer or float number but is {type(value)}"
            )
    if np.isnan(value):
        raise ValueError(f"'{param_name}' should be a non-NaN float number")
    return value


def validate_positive_float(value, param_name, optional=False):
    """``validation.py:150-196``"""
    if value is None and optional:
        return None
    value = _unwrap_scalar(value)
    try:
        value = float(value)
    except (TypeError, ValueError):
        raise ValueError(f"'{param_name}' should be a float number but is {type(value)}")
    if value <= 0:
        raise ValueError(f"'{param_name}' should be a positive float number")
    if np.isnan(value):
        raise ValueError(f"'{param_name}' should be a non-NaN float number")
    return value


def validate_float(value, param_name, optional=False):
    """``validation.py:199-250``"""
    if value is None:
        if optional:
            return None
        raise ValueError(f"'{param_name}' is None, but is required to be a float number")
    if isinstance(value, np.ndarray) and value.size == 1:
        value = np.squeeze(value)
    value = _unwrap_scalar(value)
    if not isinstance(value, (float, int)):
        try:
            value = float(value)
        except TypeError:
            raise ValueError(f"'{param_name}' should be a float number but is {type(value)}")
    if np.isnan(value):
        raise ValueError(f"'{param_name}' should be a non-NaN float number")
    return value


def validate_positive_int(value, param_name, optional=False):
    """``validation.py:253-299`` — non-negative ints pass (0 means "no landmarks")."""
    if optional and value is None:
        return None
    value = _unwrap_scalar(value)
    if not isinstance(value, int) or value < 0:
        raise ValueError(f"'{param_name}' should be a positive integer number")
    return value


def validate_bool(value, name, optional=False):
    """``validation.py:364-393``"""
    if value is None:
        if optional:
            return None
        raise TypeError(f"'{name}' can't be None.")
    if not isinstance(value, bool):
        raise TypeError(f"{name} should be of type bool, got {type(value)} instead.")
    return value


def validate_string(value, name, choices=None):
    """``validation.py:396-435``"""
    if not isinstance(value, str):
        raise TypeError(f"{name} should be of type str, got {type(value)} instead.")
    if choices and value not in choices:
        raise ValueError(f"{name} should be one of {choices}, got '{value}' instead.")
    return value


def validate_float_or_iterable_numerical(value, name, optional=False, positive=False):
    """``validation.py:438-498``"""
    if value is None and optional:
        return None
    value = _unwrap_scalar(value, only_0d=True)
    if isinstance(value, (int, float)):
        value = float(value)
        if positive and value < 0:
            raise ValueError(f"{name} should be a non-negative number or array")
        return value
    if isinstance(value, Iterable) and not isinstance(value, str):
        result = np.asarray(value, dtype=float)
        if positive and (result < 0).any():
            raise ValueError(f"All elements in {name} should be non-negative")
        return result
    raise TypeError(f"{name} should be of type int, float or iterable, got {type(value)} instead.")


def validate_1d(x):
    """``validation.py:501-525``"""
    x = np.asarray(x, dtype=float)
    if x.ndim == 0:
        x = x[None]
    if x.ndim != 1:
        raise ValueError("`x` must be exactly 1-dimensional.")
    return x


def validate_nn_distances(nn_distances, optional=False):
    """``validation.py:528-592`` — NaN / inf / <= 0 become the smallest positive distance."""
    if nn_distances is None and optional:
        return None
    if nn_distances is None:
        message = "nn_distances are required but None is given."
        logger.error(message)
        raise ValueError(message)
    nn = np.asarray(nn_distances, dtype=float)
    nan_mask, inf_mask = np.isnan(nn), np.isinf(nn)
    with np.errstate(invalid="ignore"):
        non_positive = nn <= 0
    nan_count, inf_count, neg_count = int(nan_mask.sum()), int(inf_mask.sum()), int(non_positive.sum())
    total_invalid = nan_count + inf_count + neg_count
    bad = nan_mask | inf_mask | non_positive
    detail = f"{nan_count:,} NaN, {inf_count:,} infinite, {neg_count:,} less or equal 0. "
    tail = "Please check the input data. Setting invalid distances to the minimum positive value found."
    if bad.size and bad.all():
        message = (
            f"All {total_invalid:,} computed nearest neighbor distances "
            "(`nn_distances` attribute) contain invalid values: " + detail + tail
        )
        logger.error(message)
        raise ValueError(message)
    if total_invalid == 0:
        return nn_distances if isinstance(nn_distances, np.ndarray) and nn_distances.dtype == float else nn
    fixed = np.where(~bad, nn, nn[~bad].min())
    logger.warning(
        "The computed nearest neighbor distances (`nn_distances` attribute) contain "
        f"{total_invalid:,} invalid values: " + detail + tail
    )
    return fixed


def validate_k(k, n_samples):
    """``validation.py:595-611``"""
    if not isinstance(k, int):
        message = f"Parameter k must be an integer, got {type(k).__name__} instead."
        logger.error(message)
        raise ValueError(message)
    if k < 1:
        message = f"Parameter k must be at least 1, got {k}."
        logger.error(message)
        raise ValueError(message)
    if k >= n_samples:
        message = (
            "Parameter k must be smaller than the number of samples. "
            f"Got k={k:,} with {n_samples:,} samples."
        )
        logger.error(message)
        raise ValueError(message)
