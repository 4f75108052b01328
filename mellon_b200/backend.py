"""Device backend: the host-side object every numeric call of the package goes through.

``CudaBackend`` drives ONE B200 through the C ABI (``_native``).  With ``WORLD_SIZE > 1``
(one process per GPU, e.g. under ``torchrun``) it attaches an NCCL communicator, shards the
cell axis in contiguous row blocks and lets the library all-reduce the Gram matrix, the
(loss, gradient) vector and the Hessian diagonal on the device.

There is no CPU implementation in this package.  ``set_backend`` exists so that the host
logic (estimator pipeline, validation, sharding arithmetic) can be unit-tested on a box
without a GPU against a test double that lives under ``tests/``.
"""

from __future__ import annotations

import contextlib
import ctypes as C
import logging
import os
from collections import namedtuple

import numpy as np

from . import _native as nat

logger = logging.getLogger("mellon")

_backend = None


def set_backend(backend):
    """Install a backend object (tests only); ``None`` restores lazy creation of CudaBackend."""
    global _backend
    _backend = backend


def get_backend():
    global _backend
    if _backend is None:
        _backend = CudaBackend.from_environment()
    return _backend


N_CHUNKS = 32  # leaves of the library's fixed reduction tree over the cell axis (MB_NCHUNK)


def row_block(n, rank, world):
    """Contiguous row block ``(lo, hi, per)`` of rank ``rank``: whole chunks of ceil(n / 32) rows, chunks
    [32 rank / world, 32 (rank + 1) / world) — the same rule as ``mb_row_block`` in the C ABI, so that rank
    boundaries are leaves of the library's reduction tree and every sum over cells has the same bits for any
    number of ranks.  ``per`` is the largest block of any rank (padding unit of the row gathers)."""
    world = max(int(world), 1)
    cr = max(1, -(-n // N_CHUNKS))
    lo = min(n, (rank * N_CHUNKS // world) * cr)
    hi = min(n, ((rank + 1) * N_CHUNKS // world) * cr)
    per = max(1, -(-N_CHUNKS // world)) * cr
    return lo, hi, per


class DeviceArray:
    """Handle to a float64 matrix resident on the GPU.

    ``shape`` is the GLOBAL shape.  When ``sharded`` the device holds only this rank's row
    block ``[row_lo, row_hi)``.  Converting to NumPy (``np.asarray(a)`` / ``a.numpy()``)
    downloads (and, when sharded, all-gathers) the full matrix.
    """

    __array_priority__ = 100

    def __init__(self, backend, handle, local_shape, global_rows=None, row_lo=0, vector=False):
        self._backend = backend
        self._h = handle
        self._local = tuple(int(v) for v in local_shape)
        self.sharded = global_rows is not None
        self._rows = int(global_rows) if global_rows is not None else self._local[0]
        self.row_lo = int(row_lo)
        self._vector = vector

    # -- array-like surface ---------------------------------------------------------------
    @property
    def shape(self):
        return (self._rows,) if self._vector else (self._rows, self._local[1])

    @property
    def local_shape(self):
        return self._local

    @property
    def ndim(self):
        return 1 if self._vector else 2

    dtype = np.dtype(np.float64)

    @property
    def size(self):
        return int(np.prod(self.shape))

    def __len__(self):
        return self._rows

    def numpy(self):
        return self._backend.download(self)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return self.numpy()[idx]

    @property
    def T(self):
        return self.numpy().T

    def dot(self, other):
        return self.numpy().dot(np.asarray(other))

    def __matmul__(self, other):
        return self.numpy() @ np.asarray(other)

    def __repr__(self):
        kind = "sharded " if self.sharded else ""
        return f"<DeviceArray {kind}{' x '.join(str(s) for s in self.shape)} float64 on cuda>"

    def free(self):
        if self._h is not None:
            self._backend._free(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.free()
        except Exception:
            pass


ObjectiveState = namedtuple("ObjectiveState", "L V sum_vdr mu k")


class CudaBackend:
    """The product backend: every method is a thin call into ``libmellon_b200.so``."""

    name = "cuda"

    def __init__(self, device=0, rank=0, world=1, unique_id=None):
        self.lib = nat.load_library()
        if nat.device_count() <= 0:
            raise nat.DeviceError(
                "mellon_b200 needs an sm_100a GPU (B200): no CUDA device is visible and there "
                "is no CPU fallback. " + nat.last_error()
            )
        h = C.c_void_p()
        nat.check(self.lib.mb_ctx_create(int(device), C.byref(h)), "mb_ctx_create")
        self.ctx = h
        self.device = int(device)
        self.rank, self.world = int(rank), int(world)
        if world > 1:
            if unique_id is None:
                raise ValueError("world > 1 needs the 128-byte NCCL unique id of rank 0")
            nat.check(self.lib.mb_comm_init(self.ctx, bytes(unique_id), rank, world), "mb_comm_init")
        self._progs = {}
        # host<->device traffic of the calls made through this object (bench.py's e2e accounting)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def from_environment(cls):
        """Single GPU by default; under torchrun (RANK / WORLD_SIZE / LOCAL_RANK set) one rank
        per GPU, the NCCL unique id travelling through torch.distributed (plumbing only)."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", os.environ.get("MELLON_B200_DEVICE", "0")))
        if world <= 1:
            return cls(device=local)
        from .distributed import exchange_unique_id

        uid = exchange_unique_id(rank, world)
        return cls(device=local, rank=rank, world=world, unique_id=uid)

    def close(self):
        if self.ctx is not None:
            self.lib.mb_ctx_destroy(self.ctx)
            self.ctx = None

    @contextlib.contextmanager
    def replicated(self):
        """Inside the block this rank works ALONE on whole matrices even though a communicator is attached:
        nothing is sharded and the library is told to work solo (``mb_comm_solo``), so every matrix is the whole matrix
        and every path is the one a single-GPU process takes.  Used to
        reproduce the one-GPU result inside a multi-GPU job (bench.py / tools/check_multi_gpu.py compare its bits
        with the sharded result: the reduction tree makes them identical)."""
        saved = (self.rank, self.world)
        nat.check(self.lib.mb_comm_solo(self.ctx, 1), "mb_comm_solo")
        self.rank, self.world = 0, 1
        try:
            yield self
        finally:
            self.rank, self.world = saved
            nat.check(self.lib.mb_comm_solo(self.ctx, 0), "mb_comm_solo")

    # -- bookkeeping --------------------------------------------------------------------------
    def sync(self):
        nat.check(self.lib.mb_ctx_sync(self.ctx), "mb_ctx_sync")

    def launch_count(self):
        return int(self.lib.mb_ctx_launch_count(self.ctx))

    def info(self):
        dev, sm, free, total = C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
        nat.check(self.lib.mb_ctx_info(self.ctx, C.byref(dev), C.byref(sm), C.byref(free), C.byref(total)))
        return {"device": dev.value, "n_sm": sm.value, "free_bytes": free.value, "total_bytes": total.value}

    def timer_start(self, slot=0):
        nat.check(self.lib.mb_timer_start(self.ctx, slot))

    def timer_stop(self, slot=0):
        ms = C.c_double()
        nat.check(self.lib.mb_timer_stop(self.ctx, slot, C.byref(ms)))
        return ms.value

    PROF_CLASSES = {"cov": 0, "matvec": 1, "gemm": 2, "lossgrad": 3, "other": 4, "gemm_i8": 5, "eigh": 6}

    def prof_enable(self, on=True):
        nat.check(self.lib.mb_prof_enable(self.ctx, int(bool(on))), "mb_prof_enable")

    def prof_reset(self):
        nat.check(self.lib.mb_prof_reset(self.ctx), "mb_prof_reset")

    def prof_read(self):
        """{class: (launches, device ms, algorithmic bytes or flops)} since the last reset."""
        out = {}
        for name, cls in self.PROF_CLASSES.items():
            n, ms, work = C.c_int64(), C.c_double(), C.c_double()
            nat.check(self.lib.mb_prof_read(self.ctx, cls, C.byref(n), C.byref(ms), C.byref(work)), "mb_prof_read")
            out[name] = (n.value, ms.value, work.value)
        return out

    @contextlib.contextmanager
    def range(self, name):
        """NVTX range around a host-side stage (shows up next to the library's own per-entry-point ranges)."""
        self.lib.mb_range_push(("mellon_b200: " + name).encode())
        try:
            yield
        finally:
            self.lib.mb_range_pop()

    def flush_l2(self):
        nat.check(self.lib.mb_flush_l2(self.ctx))

    def set_option(self, key, value):
        nat.check(self.lib.mb_set_option(self.ctx, key.encode(), int(value)))

    def pinned_empty(self, shape):
        """NumPy float64 array backed by page-locked host memory."""
        n = int(np.prod(shape))
        p = C.c_void_p()
        nat.check(self.lib.mb_host_alloc(max(n, 1) * 8, C.byref(p)), "mb_host_alloc")
        buf = (C.c_double * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
        self._pinned_keep = getattr(self, "_pinned_keep", [])
        self._pinned_keep.append(p)
        return arr

    # -- memory --------------------------------------------------------------------------------
    def _alloc(self, rows, cols):
        h = C.c_void_p()
        nat.check(self.lib.mb_mat_alloc(self.ctx, int(rows), int(cols), C.byref(h)), "mb_mat_alloc")
        return h

    def _free(self, h):
        if self.ctx is not None:
            self.lib.mb_mat_free(self.ctx, h)

    def empty(self, rows, cols, global_rows=None, row_lo=0, vector=False):
        """Uninitialised device matrix; with ``global_rows`` it is this rank's row block of a matrix whose
        cell axis is sharded, and the library is told so (sums over its rows then span all ranks)."""
        d = DeviceArray(self, self._alloc(rows, cols), (rows, cols), global_rows, row_lo, vector)
        if global_rows is not None:
            nat.check(self.lib.mb_mat_set_shard(d._h, int(global_rows), int(row_lo)), "mb_mat_set_shard")
        return d

    def upload(self, a, sharded=False):
        """Copy a host array to the device.  ``sharded`` keeps only this rank's row block."""
        if isinstance(a, DeviceArray):
            return a
        a = np.asarray(a, dtype=np.float64)
        vector = a.ndim == 1
        a2 = a.reshape(-1, 1) if vector else a
        if a2.ndim != 2:
            raise ValueError(f"expected a vector or a matrix, got an array with {a.ndim} dimensions")
        n = a2.shape[0]
        if sharded and self.world > 1:
            lo, hi, _ = row_block(n, self.rank, self.world)
            blk = np.ascontiguousarray(a2[lo:hi])
            d = self.empty(hi - lo, a2.shape[1], global_rows=n, row_lo=lo, vector=vector)
        else:
            blk = np.ascontiguousarray(a2)
            d = self.empty(n, a2.shape[1], global_rows=n if sharded else None, vector=vector)
        if blk.size:
            nat.check(self.lib.mb_mat_upload(self.ctx, d._h, nat.ptr(blk), 0, blk.shape[0]), "mb_mat_upload")
            self.h2d_bytes += blk.nbytes
        return d

    def _download_local(self, d):
        rows, cols = d.local_shape
        out = np.empty((rows, cols), dtype=np.float64)
        if out.size:
            nat.check(self.lib.mb_mat_download(self.ctx, d._h, nat.ptr(out), 0, rows), "mb_mat_download")
            self.d2h_bytes += out.nbytes
        return out

    def download(self, d):
        if d.sharded and self.world > 1:
            out = self.gather_rows(self._download_local(d), d.shape[0])
        else:
            out = self._download_local(d)
        return out[:, 0] if d._vector else out

    def gather_rows(self, local, n):
        """All-gather row blocks (host in, host out) through a padded device all-gather."""
        local = np.asarray(local, dtype=np.float64)
        vec = local.ndim == 1
        l2 = local.reshape(-1, 1) if vec else local
        if self.world == 1:
            return local
        _, _, per = row_block(n, self.rank, self.world)
        pad = np.zeros((per, l2.shape[1]))
        pad[: l2.shape[0]] = l2
        src = self.upload(pad)
        dst = self.empty(per * self.world, l2.shape[1])
        nat.check(self.lib.mb_comm_allgather(self.ctx, src._h, dst._h), "mb_comm_allgather")
        padded = self._download_local(dst)
        # blocks are whole reduction chunks, so their sizes may differ by a chunk: cut each rank's rows out
        out = np.concatenate([padded[r * per: r * per + (hi - lo)]
                              for r in range(self.world) for lo, hi, _ in [row_block(n, r, self.world)]], axis=0)
        return out[:, 0] if vec else out

    # -- covariance programs --------------------------------------------------------------------
    def _prog(self, cov_func, n_cols, stock_root=False):
        """Compiled covariance program of ``cov_func``, cached per (object, width, STATE): covariances are
        mutable (the reference's tests assign ``cov.active_dims`` after a first evaluation), so a cached program
        is reused only while the serialised expression tree is unchanged."""
        from .base_cov import compile_covariance

        quiet, logger.disabled = logger.disabled, True   # __getstate__ warns about user classes: once is enough
        try:
            fingerprint = repr(cov_func.__getstate__())
        except Exception:  # user-defined subclasses without serialisation: never cache
            fingerprint = None
        finally:
            logger.disabled = quiet
        key = (id(cov_func), int(n_cols), bool(stock_root))
        hit = self._progs.get(key)
        if hit is not None and hit[0] is cov_func and fingerprint is not None and hit[2] == fingerprint:
            return hit[1]
        prog = compile_covariance(cov_func, int(n_cols), stock_root=stock_root)
        self._progs[key] = (cov_func, prog, fingerprint)
        return prog

    def supports(self, cov_func, n_cols):
        from .base_cov import NotCompilable

        try:
            self._prog(cov_func, n_cols)
            return True
        except NotCompilable:
            return False

    def cov(self, cov_func, x, y, sharded=False, stock_root=False):
        """K = cov_func(x, y) on the device (K1).  ``sharded``: x is the cell matrix.

        An expression that does not fit ONE device program (more than 4 leaves / 16 ops, or a user-defined kernel
        somewhere inside) is split at its root: the operands are built separately — recursively, each by K1 when it
        compiles — and combined on the device (``mb_mat_combine``).  Only a user-defined ``k`` itself runs on the
        host (its result is uploaded); no stock kernel is ever evaluated there."""
        from .base_cov import NotCompilable, is_stock_pair, Add, Mul, Pow
        from .util import select_active_dims

        xd = self.upload(x, sharded=sharded)
        yd = xd if y is x else self.upload(y)
        try:
            prog = self._prog(cov_func, xd.local_shape[1], stock_root=stock_root)
        except NotCompilable:
            if is_stock_pair(cov_func) or (stock_root and isinstance(cov_func, (Add, Mul, Pow))):
                xs = select_active_dims(self._host_rows(x, xd), cov_func.active_dims)
                ys = xs if y is x else select_active_dims(self._host_rows(y, yd), cov_func.active_dims)
                K = self.cov(cov_func.left, xs, ys, sharded=False)
                if isinstance(cov_func, Pow):
                    op, other, value = nat.OP_POW, None, float(cov_func.right)
                elif callable(cov_func.right):
                    op = nat.OP_ADD if isinstance(cov_func, Add) else nat.OP_MUL
                    other, value = self.cov(cov_func.right, xs, ys, sharded=False), 0.0
                else:
                    op = nat.OP_ADD if isinstance(cov_func, Add) else nat.OP_MUL
                    other, value = None, float(cov_func.right)
                nat.check(self.lib.mb_mat_combine(self.ctx, op, K._h, other._h if other is not None else None, value),
                          "mb_mat_combine")
                if xd.sharded:
                    K = self._mark_rows(K, xd)
                return K
            # a user-defined Covariance subclass: run the user's own `k` on this rank's rows and upload the result
            logger.warning("Covariance %r has no device program; evaluating its k() on the host.", cov_func)
            Kh = np.asarray(cov_func.k(self._host_rows(x, xd), self._host_rows(y, yd)), dtype=np.float64)
            K = self.upload(Kh)
            return self._mark_rows(K, xd) if xd.sharded else K
        K = self.empty(xd.local_shape[0], yd.local_shape[0], global_rows=xd.shape[0] if xd.sharded else None,
                       row_lo=xd.row_lo)
        nat.check(self.lib.mb_cov_build(self.ctx, C.byref(prog.struct), xd._h, yd._h, K._h), "mb_cov_build")
        return K

    def _host_rows(self, a, d):
        """Host copy of the rows of ``a`` that the device array ``d`` (its upload) holds on this rank."""
        if isinstance(a, DeviceArray):
            out = self._download_local(a)
            return out[:, 0] if a._vector else out
        a = np.asarray(a, dtype=np.float64)
        a2 = a.reshape(-1, 1) if a.ndim == 1 else a
        return a2[d.row_lo: d.row_lo + d.local_shape[0]] if d.sharded else a2

    def _mark_rows(self, K, like):
        """Re-label a locally built matrix as this rank's row block of a sharded one (rows as in ``like``)."""
        out = DeviceArray(self, K._h, K.local_shape, like.shape[0], like.row_lo)
        K._h = None
        nat.check(self.lib.mb_mat_set_shard(out._h, int(like.shape[0]), int(like.row_lo)), "mb_mat_set_shard")
        return out

    def nn_distances(self, x, return_index=False):
        """Exact nearest-neighbour distance of every row of ``x`` (brute force on the device).
        With a communicator each rank searches its row block against all points."""
        x = np.asarray(x, dtype=np.float64)
        x2 = x.reshape(-1, 1) if x.ndim == 1 else x
        n = x2.shape[0]
        alld = self.upload(x2)
        lo, hi, _ = row_block(n, self.rank, self.world) if self.world > 1 else (0, n, n)
        blk = alld if self.world == 1 else self.upload(np.ascontiguousarray(x2[lo:hi]))
        dist = self.empty(hi - lo, 1, vector=True)
        idx = np.empty(hi - lo, dtype=np.int64)
        nat.check(self.lib.mb_nn_distances(self.ctx, blk._h, alld._h, lo, dist._h, nat.ptr(idx)), "mb_nn_distances")
        out = self._download_local(dist)[:, 0]
        if self.world > 1:
            out = self.gather_rows(out, n)
            idx = self.gather_rows(idx.astype(np.float64), n).astype(np.int64)
        return (out, idx) if return_index else out

    def cov_diag(self, cov_func, x):
        xd = self.upload(x)
        prog = self._prog(cov_func, xd.local_shape[1])
        out = self.empty(xd.local_shape[0], 1, vector=True)
        nat.check(self.lib.mb_cov_diag(self.ctx, C.byref(prog.struct), xd._h, out._h), "mb_cov_diag")
        return self.download(out)

    def cov_chol(self, cov_func, xu, diag_add):
        """(Lp, info): Lp = chol(cov(xu, xu) + diag_add I); info > 0 on a non-positive pivot."""
        from .base_cov import NotCompilable

        xd = self.upload(xu)
        m = xd.local_shape[0]
        try:
            prog = self._prog(cov_func, xd.local_shape[1])
        except NotCompilable:
            K = self.cov(cov_func, xu, xu)
            nat.check(self.lib.mb_mat_add_diag(self.ctx, K._h, float(diag_add)), "mb_mat_add_diag")
            return K, self.potrf(K)
        Lp = self.empty(m, m)
        info = nat.check(self.lib.mb_cov_chol(self.ctx, C.byref(prog.struct), xd._h, float(diag_add), Lp._h),
                         "mb_cov_chol")
        return Lp, info

    def add_diag(self, A, value):
        nat.check(self.lib.mb_mat_add_diag(self.ctx, A._h, float(value)), "mb_mat_add_diag")
        return A

    def add_diag_vec(self, A, v):
        """A(i, i) += v(i) for a replicated square matrix."""
        vd = self.upload(np.asarray(v, dtype=np.float64))
        nat.check(self.lib.mb_mat_add_diag_vec(self.ctx, A._h, vd._h), "mb_mat_add_diag_vec")
        return A

    def potrf(self, A):
        return nat.check(self.lib.mb_potrf(self.ctx, A._h), "mb_potrf")

    def lowrank_standard(self, cov_func, x, xu, Lp):
        """L = cov(x, xu) Lp^-T, rows sharded like x (K1 + K3)."""
        K = self.cov(cov_func, x, xu, sharded=True)
        Lpd = self.upload(Lp)
        nat.check(self.lib.mb_trsm_right_lt(self.ctx, Lpd._h, K._h), "mb_trsm_right_lt")
        return K

    def trsm_right_lt(self, Lp, X):
        nat.check(self.lib.mb_trsm_right_lt(self.ctx, self.upload(Lp)._h, X._h), "mb_trsm_right_lt")
        return X

    def tri_solve(self, Lp, b, trans=False):
        """Lp^-1 b (or Lp^-T b).  b: host vector / matrix; returns a host array."""
        b = np.asarray(b, dtype=np.float64)
        bd = self.upload(b.copy())
        fresh = DeviceArray(self, bd._h, bd.local_shape, vector=b.ndim == 1)
        bd._h = None
        nat.check(self.lib.mb_tri_solve(self.ctx, self.upload(Lp)._h, 1 if trans else 0, fresh._h), "mb_tri_solve")
        return self.download(fresh)

    def tri_solve_dev(self, Lp, B, trans=False):
        nat.check(self.lib.mb_tri_solve(self.ctx, self.upload(Lp)._h, 1 if trans else 0, B._h), "mb_tri_solve")
        return B

    def gram(self, L):
        r = L.local_shape[1]
        G = self.empty(r, r)
        nat.check(self.lib.mb_gram(self.ctx, L._h, G._h), "mb_gram")
        return G

    def gemv_t(self, L, t):
        """L^T t summed over all ranks; ``t`` is the full (global) vector."""
        td = self.upload(np.asarray(t, dtype=np.float64), sharded=L.sharded)
        b = self.empty(L.local_shape[1], 1, vector=True)
        nat.check(self.lib.mb_gemv_t(self.ctx, L._h, td._h, b._h), "mb_gemv_t")
        return self.download(b)

    def ridge_init(self, L, target):
        """(L^T L + I)^-1 L^T target  — sklearn Ridge(alpha=1, fit_intercept=False) primal."""
        td = self.upload(np.asarray(target, dtype=np.float64), sharded=L.sharded)
        z0 = np.empty(L.local_shape[1], dtype=np.float64)
        nat.check(self.lib.mb_ridge_init(self.ctx, L._h, td._h, nat.ptr(z0)), "mb_ridge_init")
        self.d2h_bytes += z0.nbytes
        return z0

    def gemm(self, A, B, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, out=None, reduce=False):
        """``alpha op(A) op(B) + beta out``.  ``reduce``: both operands are row-sharded and contracted
        over the cell axis (op(A) = A^T) — summed over the cells of all ranks by the library."""
        A = self.upload(A)
        B = self.upload(B, sharded=reduce)
        m = A.local_shape[1] if trans_a else A.local_shape[0]
        n = B.local_shape[0] if trans_b else B.local_shape[1]
        if out is None:
            keep_rows = A.sharded and not trans_a
            out = self.empty(m, n, global_rows=A.shape[0] if keep_rows else None, row_lo=A.row_lo if keep_rows else 0)
        if reduce and not (A.sharded and B.sharded and trans_a and not trans_b):
            raise ValueError("reduce=True contracts A^T B over the cells of two row-sharded operands")
        # with sharded operands and trans_a the library sums over the cells of all ranks (fixed tree)
        nat.check(self.lib.mb_gemm(self.ctx, int(trans_a), int(trans_b), float(alpha), A._h, B._h, float(beta),
                                   out._h), "mb_gemm")
        return out

    def eye(self, n):
        out = self.empty(n, n)
        nat.check(self.lib.mb_mat_fill(self.ctx, out._h, 0.0), "mb_mat_fill")
        return self.add_diag(out, 1.0)

    def copy(self, A):
        out = self.empty(*A.local_shape, global_rows=A.shape[0] if A.sharded else None, row_lo=A.row_lo,
                         vector=A._vector)
        nat.check(self.lib.mb_mat_copy(self.ctx, A._h, out._h), "mb_mat_copy")
        return out

    def scale(self, A, s):
        nat.check(self.lib.mb_mat_scale(self.ctx, A._h, float(s)), "mb_mat_scale")
        return A

    def row_sumsq(self, A):
        """Host vector of the squared row norms of a (non-sharded or local) device matrix."""
        out = self.empty(A.local_shape[0], 1, vector=True)
        nat.check(self.lib.mb_mat_row_sumsq(self.ctx, A._h, out._h), "mb_mat_row_sumsq")
        return self._download_local(out)[:, 0]

    def allreduce(self, A):
        nat.check(self.lib.mb_comm_allreduce(self.ctx, A._h), "mb_comm_allreduce")
        return A

    def scale_cols(self, A, s):
        sd = self.upload(np.asarray(s, dtype=np.float64))
        nat.check(self.lib.mb_mat_scale_cols(self.ctx, A._h, sd._h), "mb_mat_scale_cols")
        return A

    def scale_rows(self, A, s):
        """A(i, :) *= s(i); ``s`` is the full (global) vector, cut like the rows of a sharded A."""
        sd = self.upload(np.asarray(s, dtype=np.float64), sharded=A.sharded)
        nat.check(self.lib.mb_mat_scale_rows(self.ctx, A._h, sd._h), "mb_mat_scale_rows")
        return A

    def copy_cols(self, A, c0, ncols):
        out = self.empty(A.local_shape[0], ncols, global_rows=A.shape[0] if A.sharded else None, row_lo=A.row_lo)
        nat.check(self.lib.mb_mat_copy_cols(self.ctx, A._h, int(c0), int(ncols), out._h), "mb_mat_copy_cols")
        return out

    def copy_rows(self, A, r0, nrows):
        out = self.empty(nrows, A.local_shape[1])
        nat.check(self.lib.mb_mat_copy_rows(self.ctx, A._h, int(r0), int(nrows), out._h), "mb_mat_copy_rows")
        return out

    def sqdist_min(self, xd, xnorm_d, candidates, closest_d):
        """One k-means++ seeding step (``mb_sqdist_min``): ``out[t, i] = min(closest[i], |x_i - c_t|^2)`` as a
        device matrix and the candidates' potentials ``sum_i out[t, i]`` on the host."""
        cd = self.upload(np.ascontiguousarray(candidates, dtype=np.float64))
        T = cd.local_shape[0]
        out = self.empty(T, xd.local_shape[0])
        pot = np.empty(T, dtype=np.float64)
        nat.check(self.lib.mb_sqdist_min(self.ctx, xd._h, xnorm_d._h, cd._h, closest_d._h if closest_d is not None else None,
                                         out._h, pot.ctypes.data_as(C.POINTER(C.c_double))), "mb_sqdist_min")
        self.d2h_bytes += pot.nbytes
        return out, pot

    def nearest_rows(self, xd, centers):
        """Index of the nearest row of ``centers`` for every row of the device matrix ``xd`` (exact, brute force)."""
        cd = self.upload(np.ascontiguousarray(centers, dtype=np.float64))
        n = xd.local_shape[0]
        dist = self.empty(n, 1, vector=True)
        idx = np.empty(n, dtype=np.int64)
        nat.check(self.lib.mb_nn_distances(self.ctx, xd._h, cd._h, -(n + cd.local_shape[0] + 2), dist._h, nat.ptr(idx)),
                  "mb_nn_distances")
        self.d2h_bytes += idx.nbytes
        return idx

    def transpose(self, A):
        out = self.empty(A.local_shape[1], A.local_shape[0])
        nat.check(self.lib.mb_mat_transpose(self.ctx, A._h, out._h), "mb_mat_transpose")
        return out

    def eigh(self, A):
        """Symmetric eigen-decomposition: (w ascending [host], V columns [device]).  A is consumed."""
        n = A.local_shape[0]
        w = self.empty(n, 1, vector=True)
        nat.check(self.lib.mb_syevd(self.ctx, A._h, w._h), "mb_syevd")
        return self.download(w), A

    # -- MAP objective --------------------------------------------------------------------------
    def objective(self, L, V, sum_vdr, mu, k):
        """Bundle what K5/K6 need: L (device), V (per-cell vector, sharded like L), constants."""
        Ld = self.upload(L, sharded=True)
        Vd = self.upload(np.asarray(V, dtype=np.float64), sharded=Ld.sharded)
        return ObjectiveState(Ld, Vd, float(sum_vdr), float(mu), float(k))

    def loss_grad(self, st, z):
        z = nat.as_f64(z)
        if z.shape[0] != st.L.local_shape[1]:
            raise ValueError(f"z has {z.shape[0]} entries, L has rank {st.L.local_shape[1]}")
        loss = C.c_double()
        grad = np.empty_like(z)
        nat.check(self.lib.mb_loss_grad(self.ctx, st.L._h, st.V._h, st.sum_vdr, st.mu, st.k, nat.ptr(z),
                                        C.byref(loss), nat.ptr(grad)), "mb_loss_grad")
        self.h2d_bytes += z.nbytes
        self.d2h_bytes += grad.nbytes + 8
        return loss.value, grad

    def hess_diag(self, st, z):
        z = nat.as_f64(z)
        out = np.empty_like(z)
        nat.check(self.lib.mb_hess_diag(self.ctx, st.L._h, st.V._h, st.mu, nat.ptr(z), nat.ptr(out)), "mb_hess_diag")
        self.h2d_bytes += z.nbytes
        self.d2h_bytes += out.nbytes
        return out

    def transform(self, L, z, mu):
        """f = L z + mu for ALL cells (all-gathered when L is sharded)."""
        Ld = self.upload(L, sharded=True)
        z = nat.as_f64(z)
        if z.ndim != 1 or z.shape[0] != Ld.local_shape[1]:
            raise ValueError(f"z has shape {z.shape}, L has rank {Ld.local_shape[1]}")
        f = np.empty(Ld.local_shape[0], dtype=np.float64)
        nat.check(self.lib.mb_transform(self.ctx, Ld._h, nat.ptr(z), float(mu), nat.ptr(f)), "mb_transform")
        self.h2d_bytes += z.nbytes
        self.d2h_bytes += f.nbytes
        if Ld.sharded and self.world > 1:
            return self.gather_rows(f, Ld.shape[0])
        return f

    # -- prediction -----------------------------------------------------------------------------
    def predict_mean(self, cov_func, xq, base, weights, mu):
        """mu + cov(xq, base) @ weights without materialising the covariance (K7).
        Query rows are split across ranks and gathered back when a communicator is attached."""
        from .base_cov import NotCompilable

        xq = np.asarray(xq, dtype=np.float64)
        w = np.asarray(weights, dtype=np.float64)
        vec = w.ndim == 1
        based = self.upload(base)
        wd = self.upload(w.reshape(w.shape[0], -1))
        n = xq.shape[0]
        lo, hi, _ = row_block(n, self.rank, self.world) if self.world > 1 else (0, n, n)
        blk = np.ascontiguousarray(xq[lo:hi])
        try:
            prog = self._prog(cov_func, based.local_shape[1])
        except NotCompilable:
            # no single device program: build K (split at the root, see `cov`) and multiply on the device
            out = mu + self.gemm(self.cov(cov_func, blk, based), wd).numpy()
            if self.world > 1:
                out = self.gather_rows(out, n)
            return out[:, 0] if vec else out
        out = np.empty((hi - lo, wd.local_shape[1]), dtype=np.float64)
        nat.check(self.lib.mb_predict_mean(self.ctx, C.byref(prog.struct), nat.ptr(blk), hi - lo, blk.shape[1],
                                           based._h, wd._h, float(mu), nat.ptr(out)), "mb_predict_mean")
        self.h2d_bytes += blk.nbytes
        self.d2h_bytes += out.nbytes
        if self.world > 1:
            out = self.gather_rows(out, n)
        return out[:, 0] if vec else out

    def cov_matvec_dev(self, cov_func, xq, base, weights, mu):
        """Device-resident K7 (inputs already in HBM); returns a DeviceArray."""
        xd, based = self.upload(xq), self.upload(base)
        prog = self._prog(cov_func, based.local_shape[1])
        w = weights if isinstance(weights, DeviceArray) else self.upload(np.asarray(weights).reshape(len(weights), -1))
        out = self.empty(xd.local_shape[0], w.local_shape[1])
        nat.check(self.lib.mb_cov_matvec(self.ctx, C.byref(prog.struct), xd._h, based._h, w._h, float(mu), out._h),
                  "mb_cov_matvec")
        return out
