"""The int8 digit-slice contraction planned for K1 has an executable NumPy specification in
tools/ozaki_i8_spec.py (the CUDA prototype tools/k1_i8_proto.cu implements exactly this arithmetic).  This test
keeps the specification honest: digits are balanced base-128 and reconstruct the quantised value exactly, every
integer intermediate stays inside int32, and the recombined dot product is as accurate as a float64 matmul."""

import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ozaki_i8_spec", os.path.join(os.path.dirname(HERE), "tools", "ozaki_i8_spec.py"))
ozaki = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ozaki)


def test_digits_reconstruct_the_quantised_value():
    rng = np.random.default_rng(3)
    v = rng.standard_normal((40, 50)) * np.logspace(-8, 4, 50)
    digits, E, vq = ozaki.pack(v)
    assert digits.dtype == np.int8 and digits.min() >= -64 and digits.max() <= 63
    q = sum(digits[t].astype(np.int64) * (128 ** (7 - t)) for t in range(8))
    np.testing.assert_array_equal(np.ldexp(q.astype(float), (E - 54)[:, None]), vq)
    # quantisation error: 2^-54 of the row maximum
    assert np.max(np.abs(vq - v) / np.max(np.abs(v), axis=1, keepdims=True)) <= 2.0 ** -54


def test_contraction_matches_float64_accuracy():
    rng = np.random.default_rng(4)
    for scale in (1.0, np.logspace(-6, 3, 50)):
        x = rng.standard_normal((96, 50)) * scale
        y = rng.standard_normal((80, 50)) * scale
        dx, Ex, _ = ozaki.pack(x)
        dy, Ey, _ = ozaki.pack(y)
        ref = x.astype(np.longdouble) @ y.astype(np.longdouble).T
        nx = np.linalg.norm(x, axis=1)[:, None] * np.linalg.norm(y, axis=1)[None, :]
        got36, pairs36 = ozaki.contract(dx, Ex, dy, Ey)
        got43, pairs43 = ozaki.contract(dx, Ex, dy, Ey, max_order=8)
        assert (pairs36, pairs43) == (36, 43)
        assert np.max(np.abs(got36 - ref) / nx) < 5e-15      # 36 digit pairs: the dropped ones are <= 2e-14 worst case
        assert np.max(np.abs(got43 - ref) / nx) < 5e-16      # 43 pairs: float64-matmul accuracy
