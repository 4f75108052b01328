"""Covariance base class and algebra — the ``mellon.cov`` plugin surface (``mellon/base_cov.py``).

A covariance object is a *description*: ``compile_covariance`` flattens the expression tree
(leaf kernels, ``+ * **``, nested ``active_dims``) into the postfix ``mb_kprog`` the CUDA
kernels interpret.  ``k`` / ``__call__`` / ``diag`` evaluate on the GPU and return NumPy
arrays, so user code written against ``mellon.cov`` keeps working.  A user subclass that
overrides ``k`` itself stays usable: its own ``k`` is what gets evaluated.
"""

from __future__ import annotations

import ctypes as C
import json
import logging
import sys
from abc import ABC, abstractmethod
from datetime import datetime
from importlib import import_module

import numpy as np

from . import _native as nat
from .util import deserialize, expand_to_inactive, make_serializable, select_active_dims

MELLON_NAME = __name__.split(".")[0]


def wire_module_name(module_name):
    """Module path written into serialised documents.  This package's classes are written under the reference's
    module path (``mellon.cov``, ``mellon.conditional`` — same class names there), so a file written here loads
    in stock Mellon; the loaders below resolve those names to the classes of this package."""
    head, _, rest = module_name.partition(".")
    return "mellon" + ("." + rest if rest else "") if head == MELLON_NAME else module_name
logger = logging.getLogger("mellon")


class NotCompilable(Exception):
    """The covariance expression has no device program (user-defined k, or too large)."""


class CompiledProgram:
    """Owns the ctypes arrays behind one ``mb_kprog``."""

    def __init__(self, ops, dims):
        self.n_ops = len(ops)
        self._ops = (nat.KOp * max(len(ops), 1))(*ops)
        self._dims = (C.c_int32 * max(len(dims), 1))(*dims)
        self.struct = nat.KProg(len(ops), len(dims), self._ops, self._dims)
        self.n_leaves = sum(1 for o in ops if o.op == nat.OP_LEAF)


def _resolve(cols, active_dims):
    """Columns selected by ``x[..., active_dims]`` out of the absolute columns ``cols``."""
    if active_dims is None:
        return cols
    if np.isscalar(active_dims):
        active_dims = [active_dims]
    try:
        return np.atleast_1d(cols[active_dims])
    except (IndexError, TypeError) as e:
        raise ValueError(f"active_dims {active_dims!r} do not fit an input with {len(cols)} columns") from e


def is_stock_pair(node):
    """An Add / Mul / Pow node whose ``k`` is the stock one (not overridden by a user subclass)."""
    return isinstance(node, CovariancePair) and type(node).k in (Add.k, Mul.k, Pow.k)


def compile_covariance(cov, n_cols, stock_root=False):
    """Flatten ``cov`` for inputs with ``n_cols`` columns into a :class:`CompiledProgram`.

    ``stock_root``: the call comes from the stock ``k`` of the root node itself (a user subclass that extends a
    stock kernel and calls ``super().k``), so the root is compiled as its stock kind whatever ``type(cov).k`` is."""
    ops, dims = [], []
    depth = [0, 0]  # current, max

    def push():
        depth[0] += 1
        depth[1] = max(depth[1], depth[0])

    def visit(node, cols):
        leaf_kind = getattr(type(node), "_kind", None)
        forced = stock_root and node is cov
        if leaf_kind is not None and (forced or type(node).k is _LEAF_K.get(leaf_kind)):
            sel = _resolve(cols, node.active_dims)
            all_dims = len(sel) == n_cols and np.array_equal(sel, np.arange(n_cols))
            op = nat.KOp(nat.OP_LEAF, leaf_kind, float(node.ls), float(getattr(node, "alpha", 1.0)), 0.0,
                         len(dims), -1 if all_dims else len(sel))
            if not all_dims:
                dims.extend(int(c) for c in sel)
            ops.append(op)
            push()
            return
        if is_stock_pair(node) or (forced and isinstance(node, (Add, Mul, Pow))):
            sel = _resolve(cols, node.active_dims)
            visit(node.left, sel)
            if isinstance(node, Pow):
                ops.append(nat.KOp(nat.OP_POW, 0, 0.0, 0.0, float(node.right), 0, 0))
                return
            if callable(node.right):
                visit(node.right, sel)
            else:
                ops.append(nat.KOp(nat.OP_CONST, 0, 0.0, 0.0, float(node.right), 0, 0))
                push()
            ops.append(nat.KOp(nat.OP_ADD if isinstance(node, Add) else nat.OP_MUL, 0, 0.0, 0.0, 0.0, 0, 0))
            depth[0] -= 1
            return
        raise NotCompilable(f"{type(node).__name__} has no device program")

    visit(cov, np.arange(n_cols))
    n_leaves = sum(1 for o in ops if o.op == nat.OP_LEAF)
    if len(ops) > nat.MAX_OPS or n_leaves > nat.MAX_LEAVES or depth[1] > nat.STACK_DEPTH:
        raise NotCompilable(
            f"covariance expression too large for one device program ({len(ops)} ops, {n_leaves} leaves)"
        )
    return CompiledProgram(ops, dims)


_LEAF_K = {}  # kind -> the stock ``k`` function object (filled by cov.py)


class Covariance(ABC):
    """Base covariance function (``mellon/base_cov.py:17-224``)."""

    def __init__(self, active_dims=None):
        self.active_dims = active_dims

    def __str__(self):
        return self.__repr__()

    def __repr__(self):
        args = [f"{key}={val}" for key, val in self.__dict__.items() if key != "active_dims" or val is not None]
        return self.__class__.__name__ + "(" + ", ".join(args) + ")"

    @abstractmethod
    def k(self, x, y):
        """Covariance matrix of shape (n, m) between the rows of x and y."""

    def _device_k(self, x, y):
        from .backend import get_backend

        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if x.ndim != 2 or y.ndim != 2:
            raise ValueError("covariance inputs must be 2-dimensional (samples x features)")
        return get_backend().cov(self, x, y, stock_root=True).numpy()

    def k_grad(self, x):
        """Gradient of k(x, y) w.r.t. y: callable y -> (n, m, d).  Central differences of the
        device-evaluated kernel (the reference falls back to autodiff, base_cov.py:42-66)."""
        x = np.asarray(x, dtype=np.float64)

        def k_grad(y):
            y = np.asarray(y, dtype=np.float64)
            out = np.empty(x.shape[:-1] + y.shape)
            h = 1e-6
            for j in range(y.shape[1]):
                e = np.zeros(y.shape[1])
                e[j] = h
                out[..., j] = (np.asarray(self.k(x, y + e)) - np.asarray(self.k(x, y - e))) / (2 * h)
            return out

        return k_grad

    def __call__(self, x, y):
        return self.k(x, y)

    def diag(self, x):
        """Diagonal of k(x, x) (``base_cov.py:71-93``): the kernel at distance sqrt(1e-12)."""
        from .backend import get_backend

        x = np.asarray(x, dtype=np.float64)
        be = get_backend()
        if be.supports(self, x.shape[-1]):
            return be.cov_diag(self, x)
        if is_stock_pair(self):
            # too large for one device program (or a user kernel inside): the diagonals of the operands, combined
            xs = select_active_dims(x, self.active_dims)
            right = np.asarray(self.right.diag(xs)) if callable(self.right) else self.right
            return self._combine(np.asarray(self.left.diag(xs)), right)
        # a user-defined kernel: its own k, one point at a time
        return np.array([np.asarray(self.k(x[i : i + 1], x[i : i + 1]))[0, 0] for i in range(x.shape[0])])

    def __add__(self, other):
        return Add(self, other)

    def __radd__(self, other):
        return Add(self, other)

    def __mul__(self, other):
        return Mul(self, other)

    def __rmul__(self, other):
        return Mul(self, other)

    def __pow__(self, other):
        return Pow(self, other)

    # -- serialisation (wire format of base_cov.py:114-224) ---------------------------------
    def _data_dict(self):
        return {key: make_serializable(val) for key, val in self.__dict__.items()}

    def _metadata(self, package_only=False):
        module_name = self.__class__.__module__
        clsname = self.__class__.__name__
        if module_name == "__main__" or module_name.split(".")[0] != MELLON_NAME:
            logger.warning(
                f'The covariance function "{clsname}" is not part of {MELLON_NAME}. '
                "Make sure the implementation is available for deserialization."
            )
        meta = import_module(module_name.split(".")[0]) if module_name != "__main__" else None
        return {
            "classname": clsname,
            # pairs record the package alone, leaves the full module path (base_cov.py:254 against :124)
            "module_name": wire_module_name(module_name.split(".")[0] if package_only else module_name),
            "module_version": getattr(meta, "__version__", "NA"),
            "serialization_date": datetime.now().isoformat(),
            "python_version": sys.version,
        }

    def __getstate__(self):
        return {"type": "mellon.Covariance", "data": self._data_dict(), "metadata": self._metadata()}

    def __setstate__(self, state):
        for name, value in state["data"].items():
            setattr(self, name, deserialize(value))

    def to_json(self):
        return json.dumps(self.__getstate__())

    def to_dict(self):
        return self.__getstate__()

    @classmethod
    def from_json(cls, json_str):
        return cls.from_dict(json.loads(json_str))

    @classmethod
    def from_dict(cls, state):
        if not isinstance(state, dict) or state.get("type") != "mellon.Covariance":
            raise ValueError("The passed dict does not seem to define a covariance kernel.")
        clsname = state["metadata"]["classname"]
        module_name = state["metadata"]["module_name"]
        # files written by the reference name its own modules ("mellon.cov", ...): same classes here; a user class that
        # only shares its NAME with a stock kernel is imported from the module it names
        ours = str(module_name).split(".")[0] in ("mellon", "mellon_b200")
        Sub = _REGISTRY.get(clsname) if ours else None
        if Sub is None:
            Sub = getattr(import_module(module_name), clsname)
        instance = Sub.__new__(Sub)
        instance.__setstate__(state)
        return instance


class CovariancePair(Covariance):
    """Combination of two covariance functions (``base_cov.py:227-298``)."""

    def __init__(self, left, right, active_dims=None):
        super().__init__()
        self.left = left
        self.right = right
        self.active_dims = active_dims

    def _combine(self, a, b):  # pragma: no cover - overridden
        raise NotImplementedError

    def k(self, x, y):
        # one device program when the expression fits (<= 4 leaves); otherwise the backend builds the operands
        # separately (each on the device when it can, a user-defined kernel through its own k) and combines them
        # on the device (backend.cov)
        return self._device_k(x, y)

    def __getstate__(self):
        right = self.right.__getstate__() if callable(self.right) else make_serializable(self.right)
        return {
            "type": "mellon.Covariance",
            "left_data": self.left.__getstate__(),
            "right_data": right,
            "active_dims": make_serializable(self.active_dims),
            "metadata": self._metadata(package_only=True),
        }

    def __setstate__(self, state):
        if not isinstance(state, dict) or state.get("type") != "mellon.Covariance":
            raise ValueError("The passed dict does not seem to define a covariance kernel.")
        self.left = Covariance.from_dict(state["left_data"])
        rd = state["right_data"]
        if isinstance(rd, dict) and rd.get("type") == "mellon.Covariance":
            self.right = Covariance.from_dict(rd)
        else:
            self.right = deserialize(rd)
        self.active_dims = deserialize(state.get("active_dims", None))

    def _child_grads(self, x):
        x = np.asarray(x, dtype=np.float64)
        xs = select_active_dims(x, self.active_dims)
        gl = self.left.k_grad(xs)
        gr = self.right.k_grad(xs) if callable(self.right) else None
        return x, xs, gl, gr


class Add(CovariancePair):
    """``left + right`` (``base_cov.py:301-364``)."""

    def __repr__(self):
        return "(" + repr(self.left) + " + " + repr(self.right) + ")"

    def _combine(self, a, b):
        return a + b

    def k_grad(self, x):
        x, xs, gl, gr = self._child_grads(x)

        def k_grad(y):
            y = np.asarray(y, dtype=np.float64)
            ys = select_active_dims(y, self.active_dims)
            g = gl(ys) + (gr(ys) if gr is not None else 0.0)
            return expand_to_inactive(g, x.shape[:-1] + y.shape, self.active_dims)

        return k_grad


class Mul(CovariancePair):
    """``left * right`` (``base_cov.py:367-438``)."""

    def __repr__(self):
        return "(" + repr(self.left) + " * " + repr(self.right) + ")"

    def _combine(self, a, b):
        return a * b

    def k_grad(self, x):
        x, xs, gl, gr = self._child_grads(x)

        def k_grad(y):
            y = np.asarray(y, dtype=np.float64)
            ys = select_active_dims(y, self.active_dims)
            if gr is None:
                g = gl(ys) * self.right
            else:
                kl = np.asarray(self.left.k(xs, ys))[..., None]
                kr = np.asarray(self.right.k(xs, ys))[..., None]
                g = gl(ys) * kr + kl * gr(ys)
            return expand_to_inactive(g, x.shape[:-1] + y.shape, self.active_dims)

        return k_grad


class Pow(CovariancePair):
    """``left ** right`` (``base_cov.py:441-497``)."""

    def __repr__(self):
        return "(" + repr(self.left) + " ** " + repr(self.right) + ")"

    def _combine(self, a, b):
        return a ** b

    def k_grad(self, x):
        x, xs, gl, _ = self._child_grads(x)

        def k_grad(y):
            y = np.asarray(y, dtype=np.float64)
            ys = select_active_dims(y, self.active_dims)
            base = np.asarray(self.left.k(xs, ys))[..., None]
            g = self.right * (base ** (self.right - 1)) * gl(ys)
            return expand_to_inactive(g, x.shape[:-1] + y.shape, self.active_dims)

        return k_grad


_REGISTRY = {"Add": Add, "Mul": Mul, "Pow": Pow}
