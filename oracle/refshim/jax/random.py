"""`jax.random` restated on NumPy: the threefry2x32 counter PRNG and the samplers the reference's
tests use (PRNGKey, split, uniform, normal, multivariate_normal, choice, permutation).

Both bit layouts jax has shipped are implemented; `PARTITIONABLE` selects one (jax made
`jax_threefry_partitionable=True` the default in 0.5.0).  The reference's golden vectors
(tests/test_reference_results.py) tell which one generated them: oracle/make_golden.py tries both."""

import numpy as _np
from scipy.special import erfinv as _erfinv

from .numpy import _as_array

PARTITIONABLE = True

_U32 = _np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def _threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011), on uint32 arrays."""
    with _np.errstate(over="ignore"):
        k0, k1 = _U32(k0), _U32(k1)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = (x0 + ks[0]).astype(_U32)
        x1 = (x1 + ks[1]).astype(_U32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(i + 2) % 3] + _U32(i + 1)).astype(_U32)
    return x0, x1


def _hash_counts(key, counts):
    """jax's `threefry_2x32(keypair, count)`: split the (padded) counter array into two halves."""
    counts = _np.asarray(counts, dtype=_U32).ravel()
    n = counts.size
    if n % 2:
        counts = _np.concatenate([counts, _np.zeros(1, _U32)])
    half = counts.size // 2
    a, b = _threefry2x32(key[0], key[1], counts[:half], counts[half:])
    return _np.concatenate([a, b])[:n]


def PRNGKey(seed):
    seed = int(seed)
    return _as_array(_np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32))


key = PRNGKey


def split(key, num=2):
    key = _np.asarray(key, dtype=_U32)
    if PARTITIONABLE:
        # counts are (hi, lo) pairs of a 64-bit iota over the output shape (num,)
        a, b = _threefry2x32(key[0], key[1], _np.zeros(num, _U32), _np.arange(num, dtype=_U32))
        return _as_array(_np.stack([a, b], axis=1))
    return _as_array(_hash_counts(key, _np.arange(num * 2, dtype=_U32)).reshape(num, 2))


def _bits64(key, shape):
    key = _np.asarray(key, dtype=_U32)
    size = int(_np.prod(shape)) if len(shape) else 1
    if PARTITIONABLE:
        idx = _np.arange(size, dtype=_np.uint64)
        a, b = _threefry2x32(key[0], key[1], (idx >> _np.uint64(32)).astype(_U32), idx.astype(_U32))
        out = (a.astype(_np.uint64) << _np.uint64(32)) | b.astype(_np.uint64)
    else:
        bits = _hash_counts(key, _np.arange(2 * size, dtype=_U32))
        out = (bits[:size].astype(_np.uint64) << _np.uint64(32)) | bits[size:].astype(_np.uint64)
    return out.reshape(shape)


def _bits32(key, shape):
    key = _np.asarray(key, dtype=_U32)
    size = int(_np.prod(shape)) if len(shape) else 1
    if PARTITIONABLE:
        idx = _np.arange(size, dtype=_np.uint64)
        a, b = _threefry2x32(key[0], key[1], (idx >> _np.uint64(32)).astype(_U32), idx.astype(_U32))
        out = a ^ b
    else:
        out = _hash_counts(key, _np.arange(size, dtype=_U32))
    return out.reshape(shape)


def bits(key, shape=(), dtype=_np.uint32):
    return _as_array(_bits64(key, tuple(shape)) if _np.dtype(dtype).itemsize == 8 else _bits32(key, tuple(shape)))


def uniform(key, shape=(), dtype=float, minval=0.0, maxval=1.0):
    shape = tuple(shape) if not _np.isscalar(shape) else (shape,)
    b = _bits64(key, shape)
    one = _np.float64(1.0).view(_np.uint64)
    floats = ((b >> _np.uint64(12)) | one).view(_np.float64) - 1.0
    return _as_array(_np.maximum(minval, floats * (maxval - minval) + minval))


def normal(key, shape=(), dtype=float):
    lo = _np.nextafter(_np.float64(-1.0), _np.float64(0.0))
    u = _np.asarray(uniform(key, shape, minval=lo, maxval=1.0))
    return _as_array(_np.sqrt(2.0) * _erfinv(u))


def multivariate_normal(key, mean, cov, shape=None, dtype=float, method="cholesky"):
    mean, cov = _np.asarray(mean, dtype=float), _np.asarray(cov, dtype=float)
    shape = tuple(shape) if shape is not None else ()
    factor = _np.linalg.cholesky(cov)
    z = _np.asarray(normal(key, shape + mean.shape[-1:]))
    return _as_array(mean + _np.einsum("...ij,...j->...i", factor, z))


def permutation(key, x, axis=0, independent=False):
    """jax's sort-based shuffle: `num_rounds` passes of sorting by fresh 32-bit keys."""
    arr = _np.arange(x) if _np.isscalar(x) or _np.ndim(x) == 0 else _np.asarray(x)
    n = arr.shape[axis]
    num_rounds = int(_np.ceil(3 * _np.log(max(1, n)) / _np.log(_np.iinfo(_np.uint32).max)))
    k = _np.asarray(key, dtype=_U32)
    for _ in range(max(num_rounds, 1)):
        k, sub = (_np.asarray(s) for s in split(k))
        sort_keys = _bits32(sub, (n,))
        order = _np.argsort(sort_keys, kind="stable")
        arr = _np.take(arr, order, axis=axis)
    return _as_array(arr)


def choice(key, a, shape=(), replace=True, p=None, axis=0):
    n = int(a) if _np.ndim(a) == 0 else _np.shape(a)[axis]
    shape = tuple(shape) if not _np.isscalar(shape) else (shape,)
    size = int(_np.prod(shape)) if shape else 1
    if p is not None or replace:
        raise NotImplementedError("only uniform sampling without replacement is restated")
    idx = _np.asarray(permutation(key, n))[:size].reshape(shape)
    return _as_array(idx if _np.ndim(a) == 0 else _np.take(_np.asarray(a), idx, axis=axis))
