"""Numerical go / no-go for moving the two N*M^2 float64 GEMM families of the path (TRSM update of
L = K_NM Lp^-T, and the Gram matrix L^T L of the Ridge start) onto int8 tcgen05 digit slices.
CPU only (NumPy): the digit arithmetic is emulated EXACTLY (int8 digits, exact integer group sums through float64
matmuls whose partial sums stay below 2^53, float64 Horner recombination, float64 accumulation across K chunks), so
what is measured here is what the device kernel would return.

    python tools/ozaki_gemm_spec.py [N] [M] > profiles/ozaki_gemm_spec_r01.txt

The experiment is BASELINE config 2 in small (ExpQuad, uniform cells in [0,1)^50, ls from the heuristic: K_MM + 1e-6 I
numerically singular — the harshest case the parity study found): the whole fit is run
  (a) with LAPACK float64 everywhere (the oracle),
  (b) the same with the landmarks permuted (identical mathematics, different rounding: the reference-vs-reference floor),
  (c) with the blocked TRSM's products and the Gram matrix computed by the digit-slice scheme,
and (c) - (a) is compared with (b) - (a) in the reference's own metric std(a - b) / std(b)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.linalg as sla

from oracle import mellon_oracle as O

DBITS = int(os.environ.get("OZ_DIGIT_BITS", 7))       # bits per balanced digit: 7 (digits in [-64, 63]) or 8 ([-128, 127], the full int8 range)
NDIG = int(os.environ.get("OZ_DIGITS", 8))            # digits per value
ORDER = int(os.environ.get("OZ_ORDER", NDIG - 1))     # digit pairs kept: t + u <= ORDER
BITS = DBITS * NDIG - 2                               # |q| <= 2^BITS fits NDIG balanced digits (8 x 7 bits: 54; 7 x 8 bits: 54)
BASE, HALF = 1 << DBITS, 1 << (DBITS - 1)
KCHUNK = 32768 if DBITS == 7 else 16384               # int32 accumulators: 8 pairs * 2^12 * 32768 = 2^30; 7 pairs * 2^14 * 16384 < 2^31


def pack_rows(v):
    """Per row: power-of-two scale, BITS-bit fixed point, NDIG balanced base-128 digits (most significant first)."""
    m = np.max(np.abs(v), axis=1)
    E = np.where(m > 0, np.floor(np.log2(np.maximum(m, 1e-300))) + 1, 0).astype(np.int64)
    q = np.rint(np.ldexp(v, (BITS - E)[:, None])).astype(np.int64)
    digits = np.empty((NDIG,) + v.shape, dtype=np.int8)
    rem = q
    for t in range(NDIG - 1, -1, -1):
        d = ((rem + HALF) % BASE) - HALF
        digits[t] = d
        rem = (rem - d) // BASE
    assert np.all(rem == 0)
    return digits, E


def ozaki_nt(A, B):
    """A (n x k) times B (m x k) transposed, contraction over the contiguous index, one scale per row of A and of B —
    the operand layout of a K-major tcgen05 int8 MMA.  Long k is cut into KCHUNK pieces whose results are added in float64
    (the device flushes its int32 accumulators at the same points)."""
    n, k = A.shape
    out = np.zeros((n, B.shape[0]))
    dA, EA = pack_rows(A)
    dB, EB = pack_rows(B)
    for k0 in range(0, k, KCHUNK):
        fa = [np.ascontiguousarray(dA[t][:, k0:k0 + KCHUNK], dtype=np.float64) for t in range(NDIG)]
        fb = [np.ascontiguousarray(dB[u][:, k0:k0 + KCHUNK].T, dtype=np.float64) for u in range(NDIG)]
        S = None
        for g in range(ORDER + 1):
            G = None
            for t in range(max(0, g - NDIG + 1), min(g, NDIG - 1) + 1):
                P = fa[t] @ fb[g - t]                    # exact: |sum| <= HALF^2 * KCHUNK <= 2^28
                G = P if G is None else G + P            # exact integer group sum (int32 on the device)
            S = G if S is None else S * float(BASE) + G  # Horner in float64 (the device folds two groups on the integer pipe first: same value)
        out += S * np.ldexp(1.0, DBITS * (2 * (NDIG - 1) - ORDER))
    return out * np.ldexp(1.0, EA - BITS)[:, None] * np.ldexp(1.0, EB - BITS)[None, :]


def gemm_f64(A, B):
    return A @ B.T


def trsm_blocked(C, Lp, nt, nb=128):
    """X = C Lp^-T by block columns with inverted diagonal blocks, every product through `nt` (the device algorithm:
    mb_chol.cu trsm_inv_rec, flattened)."""
    n, m = C.shape
    X = np.empty_like(C)
    for j0 in range(0, m, nb):
        j1 = min(m, j0 + nb)
        R = C[:, j0:j1]
        if j0:
            R = R - nt(X[:, :j0], Lp[j0:j1, :j0])                       # (n x j0) . (nb x j0)^T
        Tinv = sla.solve_triangular(Lp[j0:j1, j0:j1], np.eye(j1 - j0), lower=True)
        X[:, j0:j1] = nt(R, Tinv)                                      # R . Tinv^T
    return X


def fit(X, lm, nn, nt):
    cov = O.ExpQuad(O.compute_ls(nn))
    d = X.shape[1]
    mu = O.compute_mu(nn, d)
    Lp = O.compute_Lp(X, cov, landmarks=lm)
    C = cov(X, lm)
    if nt is None:
        L = sla.solve_triangular(Lp, C.T, lower=True).T
        G = L.T @ L
    else:
        L = trsm_blocked(C, Lp, nt)
        G = nt(np.ascontiguousarray(L.T), np.ascontiguousarray(L.T))   # contraction over the cells: one scale per column of L
    t = O.mle(nn, d) - mu
    z0 = sla.solve(G + np.eye(G.shape[0]), L.T @ t, assume_a="pos")
    res = O.minimize_lbfgsb(lambda z: O.loss_and_grad(L, nn, d, mu, z), z0)
    return dict(L=L, G=G, z0=z0, z=res.pre_transformation, nfev=res.opt_state.num_fun_eval, dens=L @ res.pre_transformation + mu)


def rs(a, b):
    dlt = a - b
    return f"rel_std {np.std(dlt) / np.std(b):.2e}  max_abs/max {np.max(np.abs(dlt)) / np.max(np.abs(b)):.2e}"


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a.isdigit()]
    n = int(args[0]) if args else 16384
    m = int(args[1]) if len(args) > 1 else 1024
    X = np.random.default_rng(0).random((n, 50))
    lm = np.ascontiguousarray(X[np.sort(np.random.default_rng(1).choice(n, m, replace=False))])
    nn = O.compute_nn_distances(X)
    print(f"config 2 in small: ExpQuad, N = {n}, M = {m}, D = 50, ls = {O.compute_ls(nn):.2f}; {NDIG} digits of {DBITS} bits, pairs t + u <= {ORDER} "
          f"({sum(1 for t in range(NDIG) for u in range(NDIG) if t + u <= ORDER)} int8 products per float64 product)", flush=True)

    # the product alone, against an 80-bit reference, on the two operand shapes
    rng = np.random.default_rng(3)
    A, B = rng.standard_normal((96, 40000)) * np.logspace(0, -6, 96)[:, None], rng.standard_normal((80, 40000)) * np.logspace(-3, 0, 80)[:, None]
    ref = A.astype(np.longdouble) @ B.astype(np.longdouble).T
    bound = np.abs(A) @ np.abs(B).T
    print(f"product, k = 40000, rows scaled over 6 decades: max |err| / (|A||B|^T) digit slices {np.max(np.abs(ozaki_nt(A, B) - ref) / bound):.2e}   float64 {np.max(np.abs(A @ B.T - ref) / bound):.2e}", flush=True)

    t0 = time.time()
    a = fit(X, lm, nn, None)
    print(f"(a) float64 LAPACK: nfev {a['nfev']}  [{time.time() - t0:.0f} s]", flush=True)
    perm = np.random.default_rng(5).permutation(m)
    b = fit(X, np.ascontiguousarray(lm[perm]), nn, None)
    print(f"(b) landmarks permuted (reference-vs-reference floor): nfev {b['nfev']}   log density {rs(b['dens'], a['dens'])}", flush=True)
    t0 = time.time()
    f = fit(X, lm, nn, gemm_f64)
    print(f"(f) blocked TRSM + Gram, float64 products:            nfev {f['nfev']}   log density {rs(f['dens'], a['dens'])}   "
          f"|dL| {np.max(np.abs(f['L'] - a['L'])) / np.max(np.abs(a['L'])):.2e}  |dG| {np.max(np.abs(f['G'] - a['G'])) / np.max(np.abs(a['G'])):.2e}  |dz0| {np.max(np.abs(f['z0'] - a['z0'])) / np.max(np.abs(a['z0'])):.2e}", flush=True)
    t0 = time.time()
    c = fit(X, lm, nn, ozaki_nt)
    print(f"(c) blocked TRSM + Gram, int8 digit-slice products:   nfev {c['nfev']}   log density {rs(c['dens'], a['dens'])}   "
          f"|dL| {np.max(np.abs(c['L'] - a['L'])) / np.max(np.abs(a['L'])):.2e}  |dG| {np.max(np.abs(c['G'] - a['G'])) / np.max(np.abs(a['G'])):.2e}  |dz0| {np.max(np.abs(c['z0'] - a['z0'])) / np.max(np.abs(a['z0'])):.2e}  [{time.time() - t0:.0f} s]", flush=True)
    print(f"    (c) against (f), same algorithm, float64 products: log density {rs(c['dens'], f['dens'])}", flush=True)
