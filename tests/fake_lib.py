"""TEST DOUBLE of libmellon_b200.so at the C-ABI level (NumPy/SciPy, float64).

It lets the whole host side of the package — estimator pipeline, validation, covariance-program
compilation, ctypes marshalling, row sharding and gathers — run on a box without a GPU, including
world_size-2 runs over `gloo`.  It lives under tests/ and is installed with
``mellon_b200.set_backend(FakeBackend())``; the product never imports it.  Every entry point has
the signature of include/mellon_b200.h and receives exactly the ctypes objects backend.py passes.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.linalg import solve_triangular

from mellon_b200 import _native as nat
from mellon_b200.backend import CudaBackend


def _val(h):
    return h.value if hasattr(h, "value") else int(h)


def _host(p, n):
    """View n doubles at a host pointer (c_void_p)."""
    if n == 0:
        return np.zeros(0)
    return np.ctypeslib.as_array((C.c_double * n).from_address(_val(p)))


def _prog_of(ref):
    return ref._obj if hasattr(ref, "_obj") else ref


def eval_program(prog, x, y):
    """Interpret an ``mb_kprog`` in NumPy with the arithmetic of util.py:351-366 and cov.py."""
    stack = []
    for i in range(prog.n_ops):
        o = prog.ops[i]
        if o.op == nat.OP_LEAF:
            if o.dim_cnt < 0:
                xs, ys = x, y
            else:
                dims = [prog.dims[o.dim_off + k] for k in range(o.dim_cnt)]
                xs, ys = x[:, dims], y[:, dims]
            if o.kind == nat.K_LINEAR:
                stack.append(xs @ ys.T / o.ls)
                continue
            xx = np.sum(xs * xs, axis=1)[:, None]
            yy = np.sum(ys * ys, axis=1)[None, :]
            dist = np.sqrt(np.maximum(xx - 2 * (xs @ ys.T) + yy + 1e-12, 0))
            if o.kind == 6:
                stack.append(dist)
                continue
            r = dist / o.ls
            if o.kind == nat.K_MATERN32:
                r = np.sqrt(3.0) * r
                stack.append((r + 1) * np.exp(-r))
            elif o.kind == nat.K_MATERN52:
                r = np.sqrt(5.0) * r
                stack.append((r + r * r / 3 + 1) * np.exp(-r))
            elif o.kind == nat.K_EXPQUAD:
                stack.append(np.exp(-r * r / 2))
            elif o.kind == nat.K_EXPONENTIAL:
                stack.append(np.exp(-r / 2))
            elif o.kind == nat.K_RATQUAD:
                stack.append((r * r / (2 * o.alpha) + 1) ** -o.alpha)
            else:
                raise ValueError(f"unknown kernel kind {o.kind}")
        elif o.op == nat.OP_CONST:
            stack.append(o.value)
        elif o.op == nat.OP_POW:
            stack.append(stack.pop() ** o.value)
        else:
            b, a = stack.pop(), stack.pop()
            stack.append(a + b if o.op == nat.OP_ADD else a * b)
    (out,) = stack
    return np.broadcast_to(out, (x.shape[0], y.shape[0])).copy()


class FakeLib:
    def __init__(self, rank=0, world=1):
        self.m = {}
        self.next = 1
        self.rank, self.world = rank, world
        self.err = b""
        self.launches = 0
        self.calls = []
        self.shard = {}

    # ---- helpers ------------------------------------------------------------------------------
    def A(self, h):
        return self.m[_val(h)]

    def _new(self, a):
        k = self.next
        self.next += 1
        self.m[k] = np.array(a, dtype=np.float64).reshape(a.shape if a.ndim == 2 else (-1, 1))
        return k

    def _allreduce(self, a):
        if self.world > 1:
            import torch
            import torch.distributed as dist

            t = torch.from_numpy(a)
            dist.all_reduce(t)
        return a

    # ---- the fixed reduction tree over the cell axis (mirrors csrc/mb_reduce.cu) -----------------
    NCHUNK = 32

    def mb_row_block(self, n, rank, world, lo, hi, cr):
        c = max(1, -(-n // self.NCHUNK))
        for ref, v in ((lo, min(n, (rank * self.NCHUNK // world) * c)),
                       (hi, min(n, ((rank + 1) * self.NCHUNK // world) * c)), (cr, c)):
            if ref is not None:
                ref._obj.value = v
        return 0

    def mb_mat_set_shard(self, h, global_rows, row_lo):
        a = self.A(h)
        if global_rows < 0:
            self.shard.pop(_val(h), None)
            return 0
        if row_lo < 0 or row_lo + a.shape[0] > global_rows:
            return self._fail("mb_mat_set_shard: rows outside the matrix")
        self.shard[_val(h)] = (int(global_rows), int(row_lo))
        return 0

    def _cell_sum(self, h, fn, shape):
        """Tree sum over the cells of ``fn(i0, i1)`` (the contribution of local rows [i0, i1)): 32 global
        chunks, each summed on its own, zero-padded all-reduce when the matrix is sharded, pairwise tree."""
        a = self.A(h)
        marked = _val(h) in self.shard
        G, row_lo = self.shard.get(_val(h), (a.shape[0], 0))
        sharded = marked and self.world > 1
        cr = max(1, -(-G // self.NCHUNK))
        if not sharded and (row_lo != 0 or a.shape[0] != G):
            raise AssertionError("matrix marked as a row block but no communicator is attached")
        leaves = np.zeros((self.NCHUNK,) + tuple(shape))
        for c in range(self.NCHUNK):
            i0, i1 = min(G, c * cr) - row_lo, min(G, (c + 1) * cr) - row_lo
            if i1 > i0 and i0 >= 0 and i1 <= a.shape[0]:
                leaves[c] = fn(i0, i1)
        if sharded:
            lo = min(G, (self.rank * self.NCHUNK // self.world) * cr)
            hi = min(G, ((self.rank + 1) * self.NCHUNK // self.world) * cr)
            if (lo, hi) != (row_lo, row_lo + a.shape[0]):
                raise AssertionError(f"rank {self.rank} holds rows [{row_lo}, {row_lo + a.shape[0]}) but owns [{lo}, {hi})")
            self._allreduce(leaves)
        w = self.NCHUNK
        while w > 1:
            leaves = leaves[0:w:2] + leaves[1:w:2]
            w //= 2
        return leaves[0]

    def _fail(self, msg, code=-2):
        self.err = msg.encode()
        return code

    # ---- context --------------------------------------------------------------------------------
    def mb_last_error(self):
        return self.err

    def mb_source_hash(self):
        return b"test-double"

    def mb_version(self):
        return 100

    def mb_ctx_sync(self, ctx):
        return 0

    def mb_ctx_destroy(self, ctx):
        return 0

    def mb_ctx_launch_count(self, ctx):
        return self.launches

    def mb_ctx_info(self, ctx, dev, sm, free, total):
        dev._obj.value, sm._obj.value, free._obj.value, total._obj.value = 0, 148, 1 << 37, 1 << 37
        return 0

    def mb_timer_start(self, ctx, slot):
        import time

        self._t = getattr(self, "_t", {})
        self._t[slot] = time.perf_counter()
        return 0

    def mb_timer_stop(self, ctx, slot, ms):
        import time

        ms._obj.value = (time.perf_counter() - self._t[slot]) * 1e3
        return 0

    def mb_flush_l2(self, ctx):
        return 0

    def mb_range_push(self, name):
        self.ranges = getattr(self, "ranges", [])
        self.ranges.append(name)
        return 0

    def mb_range_pop(self):
        return 0

    def mb_set_option(self, ctx, key, value):
        return 0

    def mb_prof_enable(self, ctx, on):
        return 0

    def mb_prof_reset(self, ctx):
        return 0

    def mb_prof_read(self, ctx, cls, count, ms, work):
        count._obj.value, ms._obj.value, work._obj.value = 0, 0.0, 0.0
        return 0

    def mb_comm_solo(self, ctx, on):
        if on:
            self._saved_world = (self.rank, self.world)
            self.rank, self.world = 0, 1
        else:
            self.rank, self.world = getattr(self, "_saved_world", (self.rank, self.world))
        return 0

    def mb_comm_info(self, ctx, rank, world):
        rank._obj.value, world._obj.value = self.rank, self.world
        return 0

    def mb_comm_allreduce(self, ctx, a):
        self._allreduce(self.A(a))
        return 0

    def mb_comm_allgather(self, ctx, a, out):
        src, dst = self.A(a), self.A(out)
        if self.world == 1:
            dst[...] = src
            return 0
        import torch
        import torch.distributed as dist

        parts = [torch.empty(src.shape, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(src)))
        dst[...] = np.concatenate([p.numpy() for p in parts], axis=0)
        return 0

    # ---- matrices -------------------------------------------------------------------------------
    def mb_mat_alloc(self, ctx, rows, cols, out):
        if rows < 0 or cols < 0:
            return self._fail("negative shape")
        out._obj.value = self._new(np.full((rows, cols), np.nan))
        return 0

    def mb_mat_free(self, ctx, h):
        self.m.pop(_val(h), None)
        self.shard.pop(_val(h), None)
        return 0

    def mb_mat_upload(self, ctx, h, host, row0, nrows):
        a = self.A(h)
        a[row0:row0 + nrows] = _host(host, nrows * a.shape[1]).reshape(nrows, a.shape[1])
        return 0

    def mb_mat_download(self, ctx, h, host, row0, nrows):
        a = self.A(h)
        _host(host, nrows * a.shape[1])[:] = a[row0:row0 + nrows].ravel()
        return 0

    def mb_mat_copy(self, ctx, src, dst):
        self.A(dst)[...] = self.A(src)
        return 0

    def mb_mat_fill(self, ctx, h, v):
        self.A(h)[...] = v
        return 0

    def mb_mat_transpose(self, ctx, src, dst):
        self.A(dst)[...] = self.A(src).T
        return 0

    def mb_mat_add_diag(self, ctx, h, v):
        a = self.A(h)
        a[np.diag_indices(a.shape[0])] += v
        return 0

    def mb_mat_add_diag_vec(self, ctx, h, v):
        a = self.A(h)
        a[np.diag_indices(a.shape[0])] += self.A(v).ravel()
        return 0

    def mb_mat_scale_rows(self, ctx, h, s):
        self.A(h)[...] *= self.A(s).ravel()[:, None]
        return 0

    def mb_mat_scale_cols(self, ctx, h, s):
        self.A(h)[...] *= self.A(s).ravel()[None, :]
        return 0

    def mb_mat_copy_cols(self, ctx, src, c0, ncols, dst):
        self.A(dst)[...] = self.A(src)[:, c0:c0 + ncols]
        return 0

    def mb_mat_copy_rows(self, ctx, src, r0, nrows, dst):
        self.A(dst)[...] = self.A(src)[r0:r0 + nrows]
        return 0

    def mb_mat_symmetrize(self, ctx, h):
        a = self.A(h)
        a[...] = np.tril(a) + np.tril(a, -1).T
        return 0

    def mb_mat_scale(self, ctx, h, s):
        self.A(h)[...] *= s
        return 0

    def mb_mat_combine(self, ctx, op, a, b, value):
        x = self.A(a)
        y = self.A(b) if b is not None and _val(b) else value
        if op == nat.OP_ADD:
            x[...] = x + y
        elif op == nat.OP_MUL:
            x[...] = x * y
        elif op == nat.OP_POW:
            x[...] = x ** value
        else:
            return self._fail("mb_mat_combine: bad op")
        return 0

    def mb_mat_row_sumsq(self, ctx, a, out):
        self.A(out)[:, 0] = np.sum(self.A(a) ** 2, axis=1)
        return 0

    # ---- covariance -----------------------------------------------------------------------------
    def mb_cov_build(self, ctx, prog, x, y, K):
        self.launches += 1
        self.calls.append("mb_cov_build")
        self.A(K)[...] = eval_program(_prog_of(prog), self.A(x), self.A(y))
        return 0

    def mb_cov_diag(self, ctx, prog, x, out):
        xa = self.A(x)
        p = _prog_of(prog)
        self.A(out)[:, 0] = [eval_program(p, xa[i:i + 1], xa[i:i + 1])[0, 0] for i in range(xa.shape[0])]
        return 0

    def mb_cov_matvec(self, ctx, prog, xq, base, w, mu, out):
        self.launches += 1
        self.A(out)[...] = mu + eval_program(_prog_of(prog), self.A(xq), self.A(base)) @ self.A(w)
        return 0

    def mb_predict_mean(self, ctx, prog, xq, nq, d, base, w, mu, out):
        self.launches += 1
        self.calls.append("mb_predict_mean")
        b = self.A(base)
        if d != b.shape[1]:
            return self._fail("feature mismatch")
        q = _host(xq, nq * d).reshape(nq, d)
        wv = self.A(w)
        _host(out, nq * wv.shape[1])[:] = (mu + eval_program(_prog_of(prog), q, b) @ wv).ravel()
        return 0

    def mb_nn_distances(self, ctx, x, allp, self_offset, dist, idx_host):
        xa, ya = self.A(x), self.A(allp)
        j = np.empty(xa.shape[0], dtype=np.int64)
        best = np.empty(xa.shape[0])
        for lo in range(0, xa.shape[0], 256):                     # row blocks: the difference tensor stays small
            blk = xa[lo:lo + 256]
            d2 = ((blk[:, None, :] - ya[None, :, :]) ** 2).sum(-1)
            rows = np.arange(blk.shape[0])
            own = rows + lo + self_offset                            # the row's own column, when it has one
            ok = (own >= 0) & (own < ya.shape[0])
            d2[rows[ok], own[ok]] = np.inf
            j[lo:lo + 256] = np.argmin(d2, axis=1)
            best[lo:lo + 256] = d2[rows, j[lo:lo + 256]]
        self.A(dist)[:, 0] = np.sqrt(best)
        if _val(idx_host):
            np.ctypeslib.as_array((C.c_int64 * xa.shape[0]).from_address(_val(idx_host)))[:] = j
        return 0

    def mb_sqdist_min(self, ctx, x, xnorm, cand, closest, out, pot):
        xa, ca, xn = self.A(x), self.A(cand), self.A(xnorm).ravel()
        d = np.maximum(xn[None, :] - 2.0 * (ca @ xa.T) + np.sum(ca * ca, axis=1)[:, None], 0.0)
        if closest is not None and _val(closest):
            d = np.minimum(d, self.A(closest).ravel()[None, :])
        self.A(out)[...] = d
        pot_arr = np.ctypeslib.as_array((C.c_double * ca.shape[0]).from_address(C.addressof(pot.contents)
                                                                                 if hasattr(pot, "contents") else _val(pot)))
        pot_arr[:] = d.sum(axis=1)
        return 0

    # ---- factorisations / solves ----------------------------------------------------------------
    def mb_potrf(self, ctx, h):
        self.launches += 1
        a = self.A(h)
        sym = np.tril(a) + np.tril(a, -1).T
        try:
            a[...] = np.linalg.cholesky(sym)
            return 0
        except np.linalg.LinAlgError:
            a[...] = np.nan
            return 1

    def mb_cov_chol(self, ctx, prog, xu, diag_add, Lp):
        self.calls.append("mb_cov_chol")
        self.mb_cov_build(ctx, prog, xu, xu, Lp)
        self.mb_mat_add_diag(ctx, Lp, diag_add)
        return self.mb_potrf(ctx, Lp)

    def mb_trsm_right_lt(self, ctx, Lp, X):
        self.launches += 1
        x = self.A(X)
        if x.size:
            x[...] = solve_triangular(self.A(Lp), x.T, lower=True).T
        return 0

    def mb_tri_solve(self, ctx, Lp, trans, B):
        b = self.A(B)
        L = self.A(Lp)
        b[...] = solve_triangular(L.T if trans else L, b, lower=not trans)
        return 0

    def mb_lowrank_standard(self, ctx, prog, x, xu, Lp, L):
        self.mb_cov_build(ctx, prog, x, xu, L)
        return self.mb_trsm_right_lt(ctx, Lp, L)

    def mb_gram(self, ctx, L, G):
        self.launches += 1
        self.calls.append("mb_gram")
        l = self.A(L)
        r = l.shape[1]
        self.A(G)[...] = self._cell_sum(L, lambda i0, i1: l[i0:i1].T @ l[i0:i1], (r, r))
        return 0

    def mb_gemv_t(self, ctx, L, t, b):
        l, tv = self.A(L), self.A(t).ravel()
        v = self._cell_sum(L, lambda i0, i1: l[i0:i1].T @ tv[i0:i1], (l.shape[1],))
        self.A(b)[...] = v.reshape(self.A(b).shape)
        return 0

    def mb_ridge_init(self, ctx, L, t, z0):
        self.calls.append("mb_ridge_init")
        l, tv = self.A(L), self.A(t).ravel()
        r = l.shape[1]
        g = self._cell_sum(L, lambda i0, i1: l[i0:i1].T @ l[i0:i1], (r, r)) + np.eye(r)
        b = self._cell_sum(L, lambda i0, i1: l[i0:i1].T @ tv[i0:i1], (r,))
        _host(z0, r)[:] = np.linalg.solve(g, b)
        return 0

    def mb_gemm(self, ctx, ta, tb, alpha, A, B, beta, Cm):
        self.launches += 1
        a, b, c = self.A(A), self.A(B), self.A(Cm)
        if ta and not tb and _val(A) in self.shard:
            if alpha != 1.0 or beta != 0.0:
                return self._fail("mb_gemm: a product contracted over sharded cells takes alpha = 1, beta = 0")
            c[...] = self._cell_sum(A, lambda i0, i1: a[i0:i1].T @ b[i0:i1], c.shape)
            return 0
        prod = alpha * ((a.T if ta else a) @ (b.T if tb else b))
        c[...] = prod + (beta * c if beta != 0.0 else 0.0)
        return 0

    # ---- objective ------------------------------------------------------------------------------
    def mb_loss_grad(self, ctx, L, V, sum_vdr, mu, k, z, loss, grad):
        self.launches += 1
        self.calls.append("mb_loss_grad")
        l, v = self.A(L), self.A(V).ravel()
        r = l.shape[1]
        zv = _host(z, r).copy()

        def chunk(i0, i1):  # everything from the chunk's own rows: the same arithmetic whoever holds them
            f = l[i0:i1] @ zv + mu
            Aexp = np.exp(f + v[i0:i1])
            return np.concatenate([l[i0:i1].T @ (Aexp - 1.0), [np.sum(f - Aexp)]])

        part = self._cell_sum(L, chunk, (r + 1,))
        _host(grad, r)[:] = zv + part[:r]
        loss._obj.value = 0.5 * float(zv @ zv) + 0.5 * k * np.log(2 * np.pi) - (part[r] + sum_vdr)
        return 0

    def mb_transform(self, ctx, L, z, mu, f):
        l = self.A(L)
        _host(f, l.shape[0])[:] = l @ _host(z, l.shape[1]) + mu
        return 0

    def mb_hess_diag(self, ctx, L, V, mu, z, diag):
        l, v = self.A(L), self.A(V).ravel()
        r = l.shape[1]
        zv = _host(z, r).copy()

        def chunk(i0, i1):
            Aexp = np.exp(l[i0:i1] @ zv + mu + v[i0:i1])
            return np.einsum("i,ij,ij->j", Aexp, l[i0:i1], l[i0:i1])

        _host(diag, r)[:] = 1.0 + self._cell_sum(L, chunk, (r,))
        return 0

    def mb_syevd(self, ctx, a, w):
        m = self.A(a)
        ev, vec = np.linalg.eigh(np.tril(m) + np.tril(m, -1).T)
        m[...] = vec
        self.A(w)[:, 0] = ev
        return 0

    def mb_host_alloc(self, nbytes, out):
        buf = (C.c_char * max(int(nbytes), 1))()
        self._keep = getattr(self, "_keep", [])
        self._keep.append(buf)
        out._obj.value = C.addressof(buf)
        return 0


class FakeBackend(CudaBackend):
    """CudaBackend with the shared library swapped for :class:`FakeLib` (host logic is untouched)."""

    name = "fake"

    def __init__(self, rank=0, world=1):
        self.lib = FakeLib(rank, world)
        self.ctx = C.c_void_p(1)
        self.device = 0
        self.rank, self.world = rank, world
        self._progs = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def close(self):
        self.ctx = None
