"""Stationary covariance kernels — the ``mellon.cov`` surface (``mellon/cov.py``).

Each class only *names* a kernel; the arithmetic (pairwise distance fused with the kernel
function) runs in ``csrc/mb_cov.cu``.  Formulas, for reference (d = distance of
``mellon/util.py:351-366``, with 1e-12 added inside the square root):

=============  ===========================================  =====================
class          k(x, y)                                      reference
=============  ===========================================  =====================
Matern32       (1 + r) exp(-r),           r = sqrt(3) d/ls  ``cov.py:62-66``
Matern52       (1 + r + r^2/3) exp(-r),   r = sqrt(5) d/ls  ``cov.py:157-161``
ExpQuad        exp(-(d/ls)^2 / 2)                           ``cov.py:255-259``
Exponential    exp(-(d/ls) / 2)                             ``cov.py:352-356``
RatQuad        ((d/ls)^2 / (2 alpha) + 1)^(-alpha)          ``cov.py:453-457``
Linear         x.y / ls                                     ``cov.py:551-556``
=============  ===========================================  =====================
"""

from __future__ import annotations

from . import _native as nat
from .base_cov import _LEAF_K, _REGISTRY, Covariance


class _Leaf(Covariance):
    _kind = None

    def __init__(self, ls=1.0, active_dims=None):
        super().__init__()
        self.ls = ls
        self.active_dims = active_dims

    def k(self, x, y):
        return self._device_k(x, y)


def _leaf(name, kind, doc):
    def k(self, x, y):
        return self._device_k(x, y)

    k.__doc__ = f"{doc}  Evaluated on the GPU; returns an (n, m) float64 array."
    cls = type(name, (_Leaf,), {"_kind": kind, "k": k, "__doc__": doc, "__module__": __name__})
    _LEAF_K[kind] = k
    _REGISTRY[name] = cls
    return cls


Matern32 = _leaf("Matern32", nat.K_MATERN32, "Matern-3/2 kernel (1 + r) exp(-r), r = sqrt(3) |x-y| / ls.")
Matern52 = _leaf("Matern52", nat.K_MATERN52,
                 "Matern-5/2 kernel (1 + r + r^2/3) exp(-r), r = sqrt(5) |x-y| / ls.")
ExpQuad = _leaf("ExpQuad", nat.K_EXPQUAD, "Exponentiated quadratic kernel exp(-|x-y|^2 / (2 ls^2)).")
Exponential = _leaf("Exponential", nat.K_EXPONENTIAL, "Exponential kernel exp(-|x-y| / (2 ls)).")
Linear = _leaf("Linear", nat.K_LINEAR, "Linear kernel x.y / ls.")


class RatQuad(_Leaf):
    """Rational quadratic kernel (|x-y|^2 / (2 alpha ls^2) + 1)^(-alpha).

    ``alpha`` is the FIRST positional argument, as in the reference (``cov.py:428``)."""

    _kind = nat.K_RATQUAD

    def __init__(self, alpha=1.0, ls=1.0, active_dims=None):
        super().__init__(ls=ls, active_dims=active_dims)
        self.alpha = alpha

    def k(self, x, y):
        return self._device_k(x, y)


_LEAF_K[nat.K_RATQUAD] = RatQuad.k
_REGISTRY["RatQuad"] = RatQuad


class _Distance(_Leaf):
    """Private leaf: the bare distance of ``mellon/util.py:351-366`` (used by util.distance)."""

    _kind = 6

    def k(self, x, y):
        return self._device_k(x, y)


_LEAF_K[6] = _Distance.k
