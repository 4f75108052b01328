"""ctypes binding of ``libmellon_b200.so`` (the C ABI declared in ``include/mellon_b200.h``).

This module is the ONLY place that touches the shared library.  There is no CPU fallback:
if the library is missing, or no sm_100 device is visible, the first call that needs the
device raises :class:`NativeLibraryError` / :class:`DeviceError`.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmellon_b200.so")


class NativeLibraryError(ImportError):
    """libmellon_b200.so is missing or lacks an expected symbol."""


class DeviceError(RuntimeError):
    """A device-side call failed (message from ``mb_last_error``)."""


class KOp(C.Structure):
    """``mb_kop`` (include/mellon_b200.h)."""

    _fields_ = [
        ("op", C.c_int32),
        ("kind", C.c_int32),
        ("ls", C.c_double),
        ("alpha", C.c_double),
        ("value", C.c_double),
        ("dim_off", C.c_int32),
        ("dim_cnt", C.c_int32),
    ]


class KProg(C.Structure):
    """``mb_kprog`` (include/mellon_b200.h)."""

    _fields_ = [
        ("n_ops", C.c_int32),
        ("n_dims", C.c_int32),
        ("ops", C.POINTER(KOp)),
        ("dims", C.POINTER(C.c_int32)),
    ]


OP_LEAF, OP_CONST, OP_ADD, OP_MUL, OP_POW = range(5)
K_MATERN32, K_MATERN52, K_EXPQUAD, K_EXPONENTIAL, K_RATQUAD, K_LINEAR = range(6)
MAX_LEAVES, MAX_OPS, STACK_DEPTH = 4, 16, 4

_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_pd = C.POINTER(C.c_double)
_pprog = C.POINTER(KProg)

# name -> (restype, argtypes); mirrors include/mellon_b200.h one to one
SIGNATURES = {
    "mb_last_error": (C.c_char_p, []),
    "mb_version": (_i, []),
    "mb_source_hash": (C.c_char_p, []),
    "mb_device_count": (_i, [C.POINTER(_i)]),
    "mb_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "mb_ctx_destroy": (_i, [_vp]),
    "mb_ctx_sync": (_i, [_vp]),
    "mb_ctx_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i64)]),
    "mb_ctx_launch_count": (_i64, [_vp]),
    "mb_timer_start": (_i, [_vp, _i]),
    "mb_timer_stop": (_i, [_vp, _i, _pd]),
    "mb_prof_enable": (_i, [_vp, _i]),
    "mb_prof_reset": (_i, [_vp]),
    "mb_prof_read": (_i, [_vp, _i, C.POINTER(_i64), _pd, _pd]),
    "mb_range_push": (_i, [C.c_char_p]),
    "mb_range_pop": (_i, []),
    "mb_flush_l2": (_i, [_vp]),
    "mb_set_option": (_i, [_vp, C.c_char_p, _i]),
    "mb_host_alloc": (_i, [_i64, C.POINTER(_vp)]),
    "mb_host_free": (_i, [_vp]),
    "mb_comm_unique_id": (_i, [C.c_char_p]),
    "mb_comm_init": (_i, [_vp, C.c_char_p, _i, _i]),
    "mb_comm_destroy": (_i, [_vp]),
    "mb_comm_solo": (_i, [_vp, _i]),
    "mb_comm_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "mb_row_block": (_i, [_i64, _i, _i, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "mb_mat_set_shard": (_i, [_vp, _i64, _i64]),
    "mb_comm_allreduce": (_i, [_vp, _vp]),
    "mb_comm_allgather": (_i, [_vp, _vp, _vp]),
    "mb_mat_alloc": (_i, [_vp, _i64, _i64, C.POINTER(_vp)]),
    "mb_mat_free": (_i, [_vp, _vp]),
    "mb_mat_shape": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "mb_mat_upload": (_i, [_vp, _vp, _vp, _i64, _i64]),
    "mb_mat_download": (_i, [_vp, _vp, _vp, _i64, _i64]),
    "mb_mat_copy": (_i, [_vp, _vp, _vp]),
    "mb_mat_fill": (_i, [_vp, _vp, _d]),
    "mb_mat_transpose": (_i, [_vp, _vp, _vp]),
    "mb_mat_add_diag": (_i, [_vp, _vp, _d]),
    "mb_mat_add_diag_vec": (_i, [_vp, _vp, _vp]),
    "mb_mat_scale_cols": (_i, [_vp, _vp, _vp]),
    "mb_mat_scale_rows": (_i, [_vp, _vp, _vp]),
    "mb_mat_copy_cols": (_i, [_vp, _vp, _i64, _i64, _vp]),
    "mb_mat_copy_rows": (_i, [_vp, _vp, _i64, _i64, _vp]),
    "mb_mat_symmetrize": (_i, [_vp, _vp]),
    "mb_mat_scale": (_i, [_vp, _vp, _d]),
    "mb_mat_combine": (_i, [_vp, _i, _vp, _vp, _d]),
    "mb_mat_row_sumsq": (_i, [_vp, _vp, _vp]),
    "mb_cov_build": (_i, [_vp, _pprog, _vp, _vp, _vp]),
    "mb_cov_diag": (_i, [_vp, _pprog, _vp, _vp]),
    "mb_nn_distances": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "mb_sqdist_min": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _pd]),
    "mb_cov_matvec": (_i, [_vp, _pprog, _vp, _vp, _vp, _d, _vp]),
    "mb_predict_mean": (_i, [_vp, _pprog, _vp, _i64, _i64, _vp, _vp, _d, _vp]),
    "mb_potrf": (_i, [_vp, _vp]),
    "mb_cov_chol": (_i, [_vp, _pprog, _vp, _d, _vp]),
    "mb_trsm_right_lt": (_i, [_vp, _vp, _vp]),
    "mb_tri_solve": (_i, [_vp, _vp, _i, _vp]),
    "mb_lowrank_standard": (_i, [_vp, _pprog, _vp, _vp, _vp, _vp]),
    "mb_gram": (_i, [_vp, _vp, _vp]),
    "mb_gemv_t": (_i, [_vp, _vp, _vp, _vp]),
    "mb_ridge_init": (_i, [_vp, _vp, _vp, _vp]),
    "mb_gemm": (_i, [_vp, _i, _i, _d, _vp, _vp, _d, _vp]),
    "mb_loss_grad": (_i, [_vp, _vp, _vp, _d, _d, _d, _vp, _pd, _vp]),
    "mb_transform": (_i, [_vp, _vp, _vp, _d, _vp]),
    "mb_hess_diag": (_i, [_vp, _vp, _vp, _d, _vp, _vp]),
    "mb_syevd": (_i, [_vp, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()

# translation units of libmellon_b200.so, in the order csrc/Makefile (SRCS) and __graft_entry__.build() hash them
SOURCES = ["mb_api.cu", "mb_gemm.cu", "mb_chol.cu", "mb_infer.cu", "mb_cov.cu", "mb_solve.cu", "mb_nccl.cu",
           "mb_reduce.cu", "mb_i8.cu", "mb_kmeans.cu", "mb_cov_i8.cu"]


def source_hash() -> str:
    """First 16 hex digits of the sha256 over the library's sources (SOURCES, the two shared headers, the ABI
    header) — what ``mb_source_hash()`` of a library built from this tree returns."""
    import hashlib

    root = os.path.dirname(os.path.abspath(__file__))
    files = [os.path.join(root, "csrc", f) for f in SOURCES + ["mb_common.cuh", "mb_math.cuh"]]
    files.append(os.path.join(os.path.dirname(root), "include", "mellon_b200.h"))
    h = hashlib.sha256()
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_library(path: str | None = None):
    """Load the shared library and bind every symbol of the header (no device needed)."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or os.environ.get("MELLON_B200_LIB", LIB_PATH)
        if not os.path.exists(p):
            raise NativeLibraryError(
                f"{p} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C mellon_b200/csrc`. mellon_b200 has no CPU fallback."
            )
        try:
            lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
        except OSError as e:  # pragma: no cover - depends on the host
            raise NativeLibraryError(f"could not load {p}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise NativeLibraryError(f"{p} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _lib = lib
        return lib


def last_error() -> str:
    return (load_library().mb_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = "") -> int:
    """Raise DeviceError on a negative return code; pass non-negative codes through."""
    if rc < 0:
        raise DeviceError(f"{what or 'mellon_b200'} failed (code {rc}): {last_error()}")
    return rc


def device_count() -> int:
    lib = load_library()
    n = C.c_int(0)
    rc = lib.mb_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)
