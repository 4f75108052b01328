"""Executable model of tools/gram_i8_proto.cu's data path (CPU, NumPy): the byte layout written by pack_kernel, the
operand addressing of a no-swizzle K-major tcgen05 descriptor (address(row, k) = start + (row / 8) SBO + (row % 8) 16 +
(k / 16) LBO + k % 16 — the convention tools/microbench_umma_i8.cu validated bit-exact on B200), the issuers' digit-pair
schedule and the float64 Horner flush, transcribed index by index from the CUDA source and compared with L^T L.

    python tools/gram_i8_layout_spec.py      # prints the error in units of |L|^T |L| and exits non-zero above 1e-14
"""
import sys

import numpy as np

import os

DB = int(os.environ.get("DIGIT_BITS", 8))          # digit width of the build being modelled (-DDIGIT_BITS)
NS, KS, TA, TB = (8 if DB == 7 else 7), 32, 128, 64
HALF, BASE = 1 << (DB - 1), 1 << DB
GROUPS = {7: [(0, 7), (1, 6), (2, 5), (3, 4)], 8: [(-1, 6), (0, 5), (1, 4), (2, 3)]}[DB]   # ISSUER_GROUPS, four warps
ASLICE, BSLICE = 2 * TA * 16, 2 * TB * 16
ABLOCK, BBLOCK = NS * ASLICE, NS * BSLICE


def pack(L, r):
    rows = L.shape[0]
    npa, nks = -(-r // TA), -(-rows // KS)
    Ad = np.zeros(npa * nks * ABLOCK, dtype=np.int8)
    Bd = np.zeros(2 * npa * nks * BBLOCK, dtype=np.int8)
    scale = np.zeros(npa * TA)
    cmax = np.max(np.abs(L), axis=0)
    for pa in range(npa):
        for ks in range(nks):
            for tid in range(256):
                col, chunk = tid & 127, tid >> 7
                j = pa * TA + col
                E = 0
                if j < r:
                    if cmax[j] > 0:
                        E = int(np.frexp(cmax[j])[1])
                    if ks == 0 and chunk == 0:
                        scale[j] = np.ldexp(1.0, E - 54)
                ab = (pa * nks + ks) * ABLOCK
                bb = ((2 * pa + (col >> 6)) * nks + ks) * BBLOCK
                for c in range(16):
                    i = ks * KS + chunk * 16 + c
                    q = 0
                    if j < r and i < rows:
                        q = int(np.rint(np.ldexp(L[i, j], 54 - E)))
                    for t in range(NS - 1, -1, -1):
                        d = ((q + HALF) & (BASE - 1)) - HALF
                        q = (q - d) >> DB
                        Ad[ab + t * ASLICE + chunk * (TA * 16) + col * 16 + c] = d
                        Bd[bb + t * BSLICE + chunk * (TB * 16) + (col & 63) * 16 + c] = d
    return Ad, Bd, scale, nks


def operand(mem, start, lbo, sbo, nrows):
    """The nrows x 32 int8 operand a K-major SWIZZLE_NONE descriptor addresses."""
    out = np.empty((nrows, KS), dtype=np.int64)
    for row in range(nrows):
        for k in range(KS):
            out[row, k] = mem[start + (row // 8) * sbo + (row % 8) * 16 + (k // 16) * lbo + (k % 16)]
    return out


def gram(L, r):
    Ad, Bd, scale, nks = pack(L, r)
    npa = -(-r // TA)
    G = np.zeros((r, r))
    for pa in range(npa):
        for pb in range(2 * npa):
            if not 64 * pb < 128 * (pa + 1):
                continue
            tmem = np.zeros((128, 512), dtype=np.int64)
            for ks in range(nks):
                sa, sb = (pa * nks + ks) * ABLOCK, (pb * nks + ks) * BBLOCK       # what the two bulk copies bring
                for g0, g1 in GROUPS:
                    for g in (g1, g0):
                        for t in range(g + 1):
                            A = operand(Ad, sa + t * ASLICE, TA * 16, 128, TA)
                            B = operand(Bd, sb + (g - t) * BSLICE, TB * 16, 128, TB)
                            tmem[:, g * TB:(g + 1) * TB] += A @ B.T
            assert np.max(np.abs(tmem)) < 2 ** 31
            for row in range(128):
                gi = pa * TA + row
                if gi >= r:
                    continue
                si = scale[gi] * 2.0 ** (DB * (NS - 1))
                for c in range(TB):
                    gj = pb * TB + c
                    if gj >= r:
                        continue
                    h = 0.0
                    for g in range(NS):
                        h = float(tmem[row, g * TB + c]) if g == 0 else h * float(BASE) + float(tmem[row, g * TB + c])
                    G[gi, gj] += h * si * scale[gj]
    return G


def pack_rows(X, expo, P, ks_lo, ks_hi, nks, Xd, scale):
    """pack_rows_kernel<P>: digits of the columns of k-steps [ks_lo, ks_hi) of a row-major operand, one scale per row."""
    rows, cols = X.shape
    for panel in range(-(-rows // P)):
        for ks in range(ks_lo, ks_hi):
            for tid in range(2 * P):
                row, chunk = tid % P, tid // P
                i = panel * P + row
                E = 0
                if i < rows:
                    E = int(expo[i])
                    scale[i] = np.ldexp(1.0, E - 54)
                blk = (panel * nks + ks) * (NS * 2 * P * 16)
                for c in range(16):
                    k = ks * KS + chunk * 16 + c
                    q = 0
                    if i < rows and k < cols:
                        q = int(np.rint(np.ldexp(X[i, k], 54 - E)))
                    for t in range(NS - 1, -1, -1):
                        d = ((q + HALF) & (BASE - 1)) - HALF
                        q = (q - d) >> DB
                        Xd[blk + t * (2 * P * 16) + chunk * (P * 16) + row * 16 + c] = d


def gemm_tile(Ad, Bd, scale_a, scale_b, nks, a_stride, b_stride, pa, pb, rows_a, rows_b, alpha, out):
    """One CTA of gram_i8_kernel in its general form: out[pa*128.., pb*64..] += alpha A B^T."""
    tmem = np.zeros((128, 512), dtype=np.int64)
    for ks in range(nks):
        sa, sb = (pa * a_stride + ks) * ABLOCK, (pb * b_stride + ks) * BBLOCK
        for g0, g1 in GROUPS:
            for g in (g1, g0):
                for t in range(g + 1):
                    A = operand(Ad, sa + t * ASLICE, TA * 16, 128, TA)
                    B = operand(Bd, sb + (g - t) * BSLICE, TB * 16, 128, TB)
                    tmem[:, g * TB:(g + 1) * TB] += A @ B.T
    assert np.max(np.abs(tmem)) < 2 ** 31
    for row in range(128):
        gi = pa * TA + row
        if gi >= rows_a:
            continue
        si = alpha * scale_a[gi] * 2.0 ** (DB * (NS - 1))
        for c in range(TB):
            gj = pb * TB + c
            if gj >= rows_b:
                continue
            h = 0.0
            for g in range(NS):
                h = float(tmem[row, g * TB + c]) if g == 0 else h * float(BASE) + float(tmem[row, g * TB + c])
            out[gi, gj] += h * si * scale_b[gj]


def trsm(C, Lp):
    """run_trsm: X <- C Lp^-T, left-looking over 128-column blocks; the update of block b is the GEMM above on the digits
    of the finished columns of X (packed block by block) and of the rows of Lp, the diagonal block its float64 inverse."""
    n, m = C.shape
    assert n % TA == 0 and m % TA == 0
    nks_total = m // KS
    X = C.copy()
    ref = np.linalg.solve(Lp, C.T).T
    ex = np.frexp(np.max(np.abs(ref), axis=1))[1] + 1
    el = np.frexp(np.max(np.abs(Lp), axis=1))[1]
    Xd = np.zeros((n // TA) * nks_total * ABLOCK, dtype=np.int8)
    Lpd = np.zeros((m // TB) * nks_total * BBLOCK, dtype=np.int8)
    sx, sl = np.zeros(n), np.zeros(m)
    pack_rows(Lp, el, TB, 0, nks_total, nks_total, Lpd, sl)
    for b in range(m // TA):
        j0 = b * TA
        if b > 0:
            for pa in range(n // TA):
                for pb in (2 * b, 2 * b + 1):
                    gemm_tile(Xd, Lpd, sx, sl, j0 // KS, nks_total, nks_total, pa, pb, n, m, -1.0, X)
        Tinv = np.linalg.inv(Lp[j0:j0 + TA, j0:j0 + TA])
        X[:, j0:j0 + TA] = X[:, j0:j0 + TA] @ np.tril(Tinv).T
        pack_rows(X, ex, TA, j0 // KS, (j0 + TA) // KS, nks_total, Xd, sx)
    return X, ref


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    n, r = 70, 150                                   # ragged in both directions: 3 k-steps (last one padded), 2 A panels
    L = (rng.standard_normal((n, 1)) + 0.3 * rng.standard_normal((n, r))) * np.logspace(0, -6, r)[None, :]
    G = gram(L, r)
    ref = L.astype(np.longdouble).T @ L.astype(np.longdouble)
    bound = np.abs(L).T @ np.abs(L)
    low = np.tril_indices(r)
    err = float(np.max(np.abs(G - ref)[low] / bound[low]))
    print(f"gram_i8 layout model: N = {n}, R = {r}: max |G - L^T L| / (|L|^T |L|) over the lower triangle = {err:.2e}")
    m = 256
    K = np.eye(m) + 0.5 * np.exp(-0.5 * ((np.arange(m)[:, None] - np.arange(m)[None, :]) / 40.0) ** 2)
    X, ref = trsm(rng.random((128, m)) - 0.3, np.linalg.cholesky(K))
    err_t = float(np.max(np.abs(X - ref)) / np.max(np.abs(ref)))
    print(f"trsm model: N = 128, M = {m}: max |X - C Lp^-T| / max |X| = {err_t:.2e}")
    sys.exit(0 if err < 1e-14 and err_t < 1e-13 else 1)
