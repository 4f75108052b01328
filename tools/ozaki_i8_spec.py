"""Executable specification (NumPy, CPU) of the int8 digit-slice contraction planned for K1:
what the pack kernel, the 36 int8 MMAs and the recombination compute, and how far the result is from the
float64 dot product the reference takes.  No GPU needed:

    python tools/ozaki_i8_spec.py > profiles/ozaki_i8_spec_r01.txt

x_i . y_j is wanted to ~1e-15 of |x||y| (the rounding the reference's own `xx - 2xy + yy` carries).
  pack:   per row, v = c * x, E = exponent with max|v| < 2^E, q = rint(v * 2^(54 - E)) (|q| < 2^54), balanced base-128
          digits q = sum_t d_t 128^(7 - t), d_t in [-64, 63], t = 0 (most significant) .. 7; k padded to 64.
  MMA:    P_tu[i, j] = sum_k d_t(x_ik) d_u(y_jk)   int8 x int8 -> int32, exact; pairs with t + u <= 7 only (36).
  groups: G_g = sum_{t + u = g} P_tu  (int32, |G_g| <= 8 * 64 * 64 * 64 = 2^21)
  fold:   H_j = G_2j * 128 + G_2j+1   (int32, < 2^29)   -- integer pipe
  Horner: S = ((H_0 * 2^14 + H_1) * 2^14 + H_2) * 2^14 + H_3   in float64 (4 conversions + 3 FMA)
  scale:  x.y = S * 128^7 * 2^(E_i - 54) * 2^(E_j - 54)   (powers of two: exact)
"""
import numpy as np

rng = np.random.default_rng(0)


def pack(v):
    """rows of v -> (digits [8, n, k] int8, exponent E [n], quantised values)."""
    m = np.max(np.abs(v), axis=1)
    E = np.where(m > 0, np.floor(np.log2(np.maximum(m, 1e-300))) + 1, 0).astype(np.int64)
    q = np.rint(np.ldexp(v, (54 - E)[:, None])).astype(np.int64)
    assert np.all(np.abs(q) <= 2 ** 54)
    vq = np.ldexp(q.astype(np.float64), (E - 54)[:, None])
    digits = np.empty((8,) + v.shape, dtype=np.int8)
    rem = q.copy()
    for t in range(7, -1, -1):                      # least significant digit first
        d = ((rem + 64) % 128) - 64
        digits[t] = d
        rem = (rem - d) // 128
    assert np.all(rem == 0), "8 balanced digits must cover |q| <= 2^54"
    return digits, E, vq


def contract(dx, Ex, dy, Ey, max_order=7):
    n, m = dx.shape[1], dy.shape[1]
    G = np.zeros((max_order + 1, n, m), dtype=np.int64)
    pairs = 0
    for t in range(8):
        for u in range(8):
            if t + u <= max_order:
                P = dx[t].astype(np.int32) @ dy[u].astype(np.int32).T   # what one pair of K=32 int8 MMAs accumulates
                assert np.max(np.abs(P)) < 2 ** 31
                G[t + u] += P
                pairs += 1
    assert np.max(np.abs(G)) <= 9 * 2 ** 18
    if max_order == 7:
        H = G[0::2] * 128 + G[1::2]                  # int32-safe: < 2^29  (integer pipe)
        assert np.max(np.abs(H)) < 2 ** 31
        S = H[0].astype(np.float64)
        for j in range(1, 4):
            S = S * 16384.0 + H[j].astype(np.float64)    # one FMA each on the device
    else:                                            # plain Horner over the groups (variant with more pairs)
        S = G[0].astype(np.float64)
        for g in range(1, max_order + 1):
            S = S * 128.0 + G[g].astype(np.float64)
    scale = np.ldexp(1.0, 7 * (14 - max_order))      # 128^(14 - max_order)
    return S * scale * np.ldexp(1.0, (Ex - 54))[:, None] * np.ldexp(1.0, (Ey - 54))[None, :], pairs


def report(name, x, y, c=1.0):
    dx, Ex, xq = pack(c * x)
    dy, Ey, yq = pack(c * y)
    got, pairs = contract(dx, Ex, dy, Ey)
    got8, pairs8 = contract(dx, Ex, dy, Ey, max_order=8)
    ref = (c * x).astype(np.longdouble) @ (c * y).astype(np.longdouble).T          # 80-bit reference
    f64 = (c * x) @ (c * y).T
    nx = np.linalg.norm(c * x, axis=1)[:, None] * np.linalg.norm(c * y, axis=1)[None, :]
    e_i8 = float(np.max(np.abs(got - ref) / nx))
    e_i8_8 = float(np.max(np.abs(got8 - ref) / nx))
    e_f64 = float(np.max(np.abs(f64 - ref) / nx))
    # squared distance of coincident points through the quantised values: exactly the quantised norm identity
    sq = (xq * xq).sum(1)[:, None] + (yq * yq).sum(1)[None, :] - 2 * got
    print(f"{name:44s} pairs {pairs:2d}: max |err| / (|x||y|) = {e_i8:.2e}   ({pairs8} pairs: {e_i8_8:.2e})   float64 matmul: {e_f64:.2e}")
    return sq


if __name__ == "__main__":
    D = 50
    x = rng.random((256, D)); y = rng.random((192, D))
    report("uniform [0,1)^50 (bench workload)", x, y, c=np.sqrt(5.0) / 38.0)
    report("standard normal", rng.standard_normal((256, D)), rng.standard_normal((192, D)))
    report("wide dynamic range (columns scaled 1e-6..1e3)", rng.standard_normal((256, D)) * np.logspace(-6, 3, D), rng.standard_normal((192, D)) * np.logspace(-6, 3, D))
    report("one dominant coordinate per row", np.eye(D)[rng.integers(0, D, 256)] * 100 + rng.standard_normal((256, D)) * 1e-3, rng.standard_normal((192, D)))
    # coincident points: squared distance must stay at rounding level of the norms (then + 1e-12 c^2 takes over)
    xs = rng.random((64, D)) * 40.0
    sq = report("coincident rows, |x|^2 ~ 2.7e4", xs, xs)
    print(f"coincident points: max |sq_ii| = {np.max(np.abs(np.diag(sq))):.2e} (float64 expansion form: ~{np.finfo(float).eps * 2.7e4 * 4:.1e})")
