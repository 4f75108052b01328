"""Digit extraction of the int8 slice GEMMs (``mellon_b200/csrc/mb_i8.cu``: ``fixed54x16`` + ``digits_of4``), restated in
NumPy: the byte trick and the ``__byte_perm`` selectors give the balanced digits of the carry loop that defines them.

The kernels turn q = rint(v 2^(54-E)) into NS = 7 balanced base-256 digits d_0 (most significant) .. d_6, each in
[-128, 127].  The definition is the carry loop (least significant first: d = ((q + 128) & 255) - 128, q = (q - d) >> 8);
the kernels take the bytes of (q + 0x0080808080808080) ^ 0x0080808080808080 instead and gather byte b of four
consecutive values into one 32-bit word with two levels of byte permutes.
"""

import numpy as np

NS = 7
BIAS = 0x0080808080808080
MASK64 = (1 << 64) - 1


def digits_loop(q):
    """The definition: balanced digits, most significant first."""
    q = int(q)
    out = [0] * NS
    for t in range(NS - 1, -1, -1):
        d = ((q + 128) & 255) - 128
        q = (q - d) >> 8
        out[t] = d
    assert q == 0
    return out


def byte_perm(x, y, s):
    """CUDA ``__byte_perm``: result byte i = byte (nibble i of s) of the 8 bytes {x: 0-3, y: 4-7}."""
    src = [(x >> (8 * i)) & 255 for i in range(4)] + [(y >> (8 * i)) & 255 for i in range(4)]
    return sum(src[(s >> (4 * i)) & 7] << (8 * i) for i in range(4))


def digits_of4(q4):
    """``digits_of4`` of mb_i8.cu: one 32-bit word per slice, value c in byte c."""
    lo, hi = [], []
    for q in q4:
        v = ((int(q) + BIAS) & MASK64) ^ BIAS
        lo.append(v & 0xFFFFFFFF)
        hi.append(v >> 32)
    w = [0] * NS
    a, b = byte_perm(lo[0], lo[1], 0x5140), byte_perm(lo[2], lo[3], 0x5140)
    w[6], w[5] = byte_perm(a, b, 0x5410), byte_perm(a, b, 0x7632)
    a, b = byte_perm(lo[0], lo[1], 0x7362), byte_perm(lo[2], lo[3], 0x7362)
    w[4], w[3] = byte_perm(a, b, 0x5410), byte_perm(a, b, 0x7632)
    a, b = byte_perm(hi[0], hi[1], 0x5140), byte_perm(hi[2], hi[3], 0x5140)
    w[2], w[1] = byte_perm(a, b, 0x5410), byte_perm(a, b, 0x7632)
    a, b = byte_perm(hi[0], hi[1], 0x7362), byte_perm(hi[2], hi[3], 0x7362)
    w[0] = byte_perm(a, b, 0x5410)
    return w


def as_int8(byte):
    return byte - 256 if byte >= 128 else byte


def test_byte_trick_gives_the_balanced_digits():
    rng = np.random.default_rng(0)
    qs = [0, 1, -1, 127, 128, -128, -129, 255, 256, 2**54, -(2**54), 2**54 - 1, -(2**54) + 1, 2**53 + 2**52, 0x7F7F7F7F7F7F7F // 2]
    qs += [int(v) for v in rng.integers(-(2**54), 2**54, size=2000, endpoint=True)]
    qs += [int(v) for v in rng.integers(-300, 300, size=200)]
    while len(qs) % 4:
        qs.append(0)
    for g in range(0, len(qs), 4):
        q4 = qs[g:g + 4]
        words = digits_of4(q4)
        for c, q in enumerate(q4):
            want = digits_loop(q)
            got = [as_int8((words[t] >> (8 * c)) & 255) for t in range(NS)]
            assert got == want, (q, got, want)
            assert sum(d * 256 ** (NS - 1 - t) for t, d in enumerate(got)) == q


def test_scaling_by_the_power_of_two_is_the_ldexp_of_the_definition():
    """``fixed54x16`` multiplies by 2^(54-E) built from its exponent bits; the definition is llrint(ldexp(v, 54-E))."""
    rng = np.random.default_rng(1)
    for scale in (1e-300, 1e-280, 1e-30, 1.0, 1e30, 1e300):
        v = rng.standard_normal(4000) * scale
        v[:4] = [0.0, -0.0, np.max(np.abs(v)), -np.max(np.abs(v))]
        v[4:8] = [5e-324, -5e-324, 2.2e-308, scale * 2.0**-60]
        m = np.max(np.abs(v))
        E = int(np.frexp(m)[1])                       # |v| < 2^E
        if E < -960:
            continue                                  # the kernels call ldexp itself there
        s = np.array([(1023 + 54 - E) << 52], dtype=np.int64).view(np.float64)[0]
        assert s == np.ldexp(1.0, 54 - E)
        fast = np.rint(v * s)
        ref = np.rint(np.ldexp(v, 54 - E))
        assert np.array_equal(fast, ref)
        assert np.max(np.abs(fast)) <= 2.0**54
