"""tools/gram_i8_layout_spec.py is the executable model of the round-2 Gram prototype (tools/gram_i8_proto.cu): byte
layout of the transposing pack, no-swizzle K-major descriptor addressing, digit-pair schedule of the four issuing
warps, float64 Horner flush — transcribed index by index from the CUDA source.  This keeps the model (and with it the
prototype's indexing) honest on ragged shapes; the kernel itself has not run on a GPU yet."""

import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("gram_i8_layout_spec",
                                              os.path.join(os.path.dirname(HERE), "tools", "gram_i8_layout_spec.py"))
model = importlib.util.module_from_spec(spec)
spec.loader.exec_module(model)


def test_layout_model_reproduces_the_gram_matrix():
    rng = np.random.default_rng(1)
    for n, r in ((33, 64), (70, 150)):            # padded last k-step; one and two 128-column panels, ragged columns
        L = (rng.standard_normal((n, 1)) + 0.3 * rng.standard_normal((n, r))) * np.logspace(0, -5, r)[None, :]
        G = model.gram(L, r)
        ref = L.astype(np.longdouble).T @ L.astype(np.longdouble)
        bound = np.abs(L).T @ np.abs(L)
        low = np.tril_indices(r)
        assert float(np.max(np.abs(G - ref)[low] / bound[low])) < 5e-15


def test_pack_writes_both_layouts_consistently():
    rng = np.random.default_rng(2)
    L = rng.standard_normal((40, 130))
    Ad, Bd, scale, nks = model.pack(L, 130)
    assert nks == 2 and Ad.dtype == np.int8 and Ad.min() >= -model.HALF and Ad.max() <= model.HALF - 1
    # column 70 of A panel 0 is column 6 of B panel 1: same 16-byte pieces in both layouts, every slice and k-step
    for ks in range(nks):
        for t in range(model.NS):
            for chunk in range(2):
                a = (0 * nks + ks) * model.ABLOCK + t * model.ASLICE + chunk * (model.TA * 16) + 70 * 16
                b = (1 * nks + ks) * model.BBLOCK + t * model.BSLICE + chunk * (model.TB * 16) + 6 * 16
                np.testing.assert_array_equal(Ad[a:a + 16], Bd[b:b + 16])
    # the digits reconstruct the quantised values
    q = np.zeros(16, dtype=np.int64)
    for t in range(model.NS):
        a = t * model.ASLICE + 70 * 16
        q += Ad[a:a + 16].astype(np.int64) * model.BASE ** (model.NS - 1 - t)
    np.testing.assert_allclose(q * scale[70], L[:16, 70], rtol=0, atol=2.0 ** -53 * np.max(np.abs(L[:, 70])))


def test_trsm_model_solves_the_triangular_system():
    """Left-looking TRSM whose block updates run through the modelled int8 GEMM on digits packed block by block."""
    rng = np.random.default_rng(3)
    m = 256
    K = np.eye(m) + 0.5 * np.exp(-0.5 * ((np.arange(m)[:, None] - np.arange(m)[None, :]) / 40.0) ** 2)
    X, ref = model.trsm(rng.random((128, m)) - 0.3, np.linalg.cholesky(K))
    assert float(np.max(np.abs(X - ref)) / np.max(np.abs(ref))) < 1e-13
