"""Argument validation with the reference's error contract (``mellon/validation.py``).

Same function names, same exception types and the same message text the reference's tests
assert on.  Arrays come back as NumPy float64; an input that already is a float64 ndarray (or
a device handle) is returned unchanged, which the estimators' ``self.x is x`` identity rule
(``mellon/base_model.py:200-204``) relies on.

The checks share three small helpers: ``_absent`` (the ``None`` / ``optional`` rule), ``_as_number`` (scalar
coercion with the caller's complaint) and ``_reject`` (log, then raise).
"""

from __future__ import annotations

import logging
from collections.abc import Iterable

import numpy as np

logger = logging.getLogger("mellon")


def _unwrap_scalar(value, only_0d=False):
    """NumPy / device scalars -> Python scalars (so isinstance checks behave)."""
    item = getattr(value, "item", None)
    if not callable(item) or (only_0d and getattr(value, "ndim", None) != 0):
        return value
    try:
        return item()
    except (ValueError, TypeError):
        return value


def _absent(value, optional, complaint=None, exc=TypeError):
    """True when ``value`` is an allowed ``None``; raises ``complaint`` for a ``None`` that is not allowed (validators
    without a complaint let the ``None`` run into their type check, as the reference's do)."""
    if value is not None:
        return False
    if optional:
        return True
    if complaint is None:
        return False
    raise exc(complaint)


def _as_number(value, complaint, catch=(TypeError,), always=False):
    """``float(value)`` for anything that is not already a Python number; ``complaint(value)`` is the error text."""
    if always or not isinstance(value, (float, int)):
        try:
            return float(value)
        except catch:
            raise ValueError(complaint(value)) from None
    return value


def _not_nan(value, param_name):
    if np.isnan(value):
        raise ValueError(f"'{param_name}' should be a non-NaN float number")
    return value


def _reject(message):
    logger.error(message)
    raise ValueError(message)


def validate_array(iterable, name, optional=False, ndim=None):
    """``validation.py:302-361`` — to float64 array; TypeError for None / non-iterables."""
    from .backend import DeviceArray

    if _absent(iterable, optional, f"'{name}' can't be None."):
        return None
    if isinstance(iterable, DeviceArray):
        array = iterable
    elif hasattr(iterable, "todense"):
        array = np.asarray(iterable.todense(), dtype=float)
    elif isinstance(iterable, Iterable):
        array = np.asarray(iterable, dtype=float)
    else:
        raise TypeError(f"'{name}' should be iterable or sparse, got {type(iterable)} instead.")
    allowed = None if ndim is None else ((ndim,) if isinstance(ndim, int) else tuple(ndim))
    if allowed is not None and array.ndim not in allowed:
        raise ValueError(f"'{name}' must be a {allowed}-dimensional array, got {array.ndim}-dimensional array instead.")
    return array


def validate_time_x(x, times=None, n_features=None, cast_scalar=False):
    """``validation.py:23-101`` — append ``times`` as the last column of ``x``."""
    x = validate_array(x, "x", ndim=2)
    n = x.shape[0]
    if cast_scalar and times is not None and (np.isscalar(times) or all(s == 1 for s in np.shape(times))):
        times = np.full(n, np.asarray(times, dtype=float).reshape(-1)[0])
    times = validate_array(times, "times", optional=True, ndim=(1, 2))
    if times is not None:
        if times.ndim == 2 and times.shape[1] != 1:
            raise ValueError("'times' must be a 1D array or a 2D array with 1 column.")
        column = times.reshape(-1, 1)
        if column.shape[0] != n:
            raise ValueError("'x' and 'times' must have the same number of samples. "
                             f"Got {n} for 'x' and {column.shape[0]} for 'times'.")
        x = np.concatenate((x, column), axis=1)
    if n_features is None:
        return x
    found = x.shape[1]
    if times is None and found == n_features - 1:
        raise ValueError(f"Expected {n_features} features including 'times' in 'x' but "
                         f"only found {found} features and 'times' is not provided.")
    if found != n_features:
        raise ValueError(f"Wrong number of features in 'x'. Expected {n_features} but got {found}.")
    return x


def validate_float_or_int(value, param_name, optional=False):
    """``validation.py:104-147``"""
    if _absent(value, optional):
        return None
    number = _as_number(_unwrap_scalar(value),
                        lambda v: f"'{param_name}' should be a positive integer or float number but is {type(v)}")
    return _not_nan(number, param_name)


def validate_positive_float(value, param_name, optional=False):
    """``validation.py:150-196``"""
    if _absent(value, optional):
        return None
    number = _as_number(_unwrap_scalar(value), lambda v: f"'{param_name}' should be a float number but is {type(v)}",
                        catch=(TypeError, ValueError), always=True)
    if number <= 0:
        raise ValueError(f"'{param_name}' should be a positive float number")
    return _not_nan(number, param_name)


def validate_float(value, param_name, optional=False):
    """``validation.py:199-250``"""
    if _absent(value, optional, f"'{param_name}' is None, but is required to be a float number", exc=ValueError):
        return None
    if isinstance(value, np.ndarray) and value.size == 1:
        value = np.squeeze(value)
    number = _as_number(_unwrap_scalar(value), lambda v: f"'{param_name}' should be a float number but is {type(v)}")
    return _not_nan(number, param_name)


def validate_positive_int(value, param_name, optional=False):
    """``validation.py:253-299`` — non-negative ints pass (0 means "no landmarks")."""
    if _absent(value, optional):
        return None
    value = _unwrap_scalar(value)
    if isinstance(value, int) and value >= 0:
        return value
    raise ValueError(f"'{param_name}' should be a positive integer number")


def _typed(value, kind, name):
    if not isinstance(value, kind):
        raise TypeError(f"{name} should be of type {kind.__name__}, got {type(value)} instead.")
    return value


def validate_bool(value, name, optional=False):
    """``validation.py:364-393``"""
    if _absent(value, optional, f"'{name}' can't be None."):
        return None
    return _typed(value, bool, name)


def validate_string(value, name, choices=None):
    """``validation.py:396-435``"""
    _typed(value, str, name)
    if choices and value not in choices:
        raise ValueError(f"{name} should be one of {choices}, got '{value}' instead.")
    return value


def validate_float_or_iterable_numerical(value, name, optional=False, positive=False):
    """``validation.py:438-498``"""
    if _absent(value, optional):
        return None
    value = _unwrap_scalar(value, only_0d=True)
    scalar = isinstance(value, (int, float))
    if not scalar and (isinstance(value, str) or not isinstance(value, Iterable)):
        raise TypeError(f"{name} should be of type int, float or iterable, got {type(value)} instead.")
    result = float(value) if scalar else np.asarray(value, dtype=float)
    if positive and np.any(np.asarray(result) < 0):
        raise ValueError(f"{name} should be a non-negative number or array" if scalar
                         else f"All elements in {name} should be non-negative")
    return result


def validate_1d(x):
    """``validation.py:501-525``"""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    if x.ndim > 1:
        raise ValueError("`x` must be exactly 1-dimensional.")
    return x


def validate_nn_distances(nn_distances, optional=False):
    """``validation.py:528-592`` — NaN / inf / <= 0 become the smallest positive distance."""
    if nn_distances is None:
        return None if optional else _reject("nn_distances are required but None is given.")
    nn = np.asarray(nn_distances, dtype=float)
    with np.errstate(invalid="ignore"):
        masks = (np.isnan(nn), np.isinf(nn), nn <= 0)
    counts = [int(m.sum()) for m in masks]
    total_invalid = sum(counts)
    if total_invalid == 0:
        return nn_distances if isinstance(nn_distances, np.ndarray) and nn_distances.dtype == float else nn
    bad = masks[0] | masks[1] | masks[2]
    detail = (f"{counts[0]:,} NaN, {counts[1]:,} infinite, {counts[2]:,} less or equal 0. "
              "Please check the input data. Setting invalid distances to the minimum positive value found.")
    if bad.all():
        _reject(f"All {total_invalid:,} computed nearest neighbor distances "
                "(`nn_distances` attribute) contain invalid values: " + detail)
    logger.warning("The computed nearest neighbor distances (`nn_distances` attribute) contain "
                   f"{total_invalid:,} invalid values: " + detail)
    return np.where(bad, nn[~bad].min(), nn)


def validate_k(k, n_samples):
    """``validation.py:595-611``"""
    if not isinstance(k, int):
        _reject(f"Parameter k must be an integer, got {type(k).__name__} instead.")
    if k < 1:
        _reject(f"Parameter k must be at least 1, got {k}.")
    if k >= n_samples:
        _reject("Parameter k must be smaller than the number of samples. "
                f"Got k={k:,} with {n_samples:,} samples.")
