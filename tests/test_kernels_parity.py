"""GPU parity, kernel by kernel: every C-ABI entry point against the CPU oracle on the same
seeded inputs (float64; tolerances stated per test, relative to the magnitude of the result)."""

import numpy as np
import pytest
from scipy.linalg import solve_triangular

import mellon_b200 as mb
from mellon_b200 import cov as C
from oracle import mellon_oracle as O



def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


PAIRS = [
    (C.Matern32, O.Matern32), (C.Matern52, O.Matern52), (C.ExpQuad, O.ExpQuad),
    (C.Exponential, O.Exponential), (C.Linear, O.Linear),
]


@pytest.mark.parametrize("n,m,d", [(1, 1, 1), (7, 5, 3), (129, 65, 10), (300, 257, 50), (64, 128, 51),
                                   (1000, 333, 2), (130, 70, 200)])
@pytest.mark.parametrize("pair", PAIRS, ids=lambda p: p[0].__name__)
@pytest.mark.parametrize("variant", [0, 1, 2], ids=["dmma", "dfma", "general"])
def test_cov_build_matches_oracle(be, variant, pair, n, m, d):
    """K1 through all three kernels: DMMA tiles + lean sqrt/exp (default), DFMA register tiles, general."""
    rng = np.random.default_rng(n * 1000 + m + d)
    x, y = rng.random((n, d)), rng.random((m, d))
    ls = 0.7 * np.sqrt(d)
    be.set_option("cov", variant)
    try:
        K = np.asarray(pair[0](ls)(x, y))
    finally:
        be.set_option("cov", 0)
    Kref = pair[1](ls)(x, y)
    assert K.shape == (n, m)
    np.testing.assert_allclose(K, Kref, rtol=0, atol=2e-13 * max(1.0, np.abs(Kref).max()))


SCALE2 = {"Matern32": 3.0, "Matern52": 5.0, "ExpQuad": 0.5, "Exponential": 0.25}  # (r / (dist / ls))^2


@pytest.mark.parametrize("pair", PAIRS[:4], ids=lambda p: p[0].__name__)
def test_cov_build_dmma_wide_range(be, pair):
    """The lean exp / sqrt of the DMMA path over many decades of r: short and long length scales, near and
    far points (k underflows towards 0), odd shapes (scalar-store tail).

    Tolerance: the expansion xx - 2xy + yy carries ~eps (|x|^2 + |y|^2) of absolute noise in the squared
    distance on BOTH sides (reference and kernel round it differently), which a kernel value turns into a
    relative error of eps c^2 (|x|^2 + |y|^2) |dlog k / d(r^2)| with |dlog k / d(r^2)| <= max(1, 1 / (2 r)).
    Allow 64x that plus 2e-12."""
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.random((150, 7)), 3.0 + rng.random((33, 7))])
    y = np.concatenate([x[:40] + 0.05, rng.random((61, 7)) * 3.0])
    sqn = (x * x).sum(1)[:, None] + (y * y).sum(1)[None, :]
    dist = O.distance(x, y)
    eps = np.finfo(float).eps
    for ls in (0.05, 1.0, 38.0, 1e4):
        K = np.asarray(pair[0](ls)(x, y))
        Kref = pair[1](ls)(x, y)
        c2 = SCALE2[pair[0].__name__] / ls ** 2
        r = np.sqrt(c2) * dist
        tol = 2e-12 + 64 * eps * c2 * sqn * np.maximum(1.0, 0.5 / r)
        err = np.abs(K - Kref) / np.maximum(np.abs(Kref), 1e-300)
        assert np.all((err <= tol) | (np.abs(K - Kref) < 1e-300)), float(np.max(err / tol))


def test_ratquad_and_alpha_first_positional(be):
    rng = np.random.default_rng(3)
    x, y = rng.random((50, 6)), rng.random((41, 6))
    K = np.asarray(C.RatQuad(2.5, 1.3)(x, y))
    np.testing.assert_allclose(K, O.RatQuad(2.5, 1.3)(x, y), rtol=0, atol=1e-13)


def test_distance_semantics(be):
    """+1e-12 inside the sqrt: d(x, x) == 1e-6 exactly-ish, and the diag of a kernel is not 1."""
    x = np.random.default_rng(0).random((20, 4))
    D = np.asarray(mb.util.distance(x, x))
    # compare SQUARED distances: on the diagonal sq = 1e-12 + (cancellation residue of xx - 2xy + yy),
    # and that residue (a few ulp of xx) depends on the summation order of the dot products
    np.testing.assert_allclose(D * D, O.distance(x, x) ** 2, rtol=0, atol=16 * np.finfo(float).eps * 4)
    assert np.allclose(np.diag(D), 1e-6, rtol=1e-3)
    dg = C.Matern52(1.0).diag(x)
    np.testing.assert_allclose(dg, O.Matern52(1.0).diag(x), rtol=0, atol=1e-15)
    assert np.all(dg < 1.0)


ACTIVE = [None, 2, -1, slice(1, 4), [0, 3], [True, False, True, False, True]]


@pytest.mark.parametrize("ad", ACTIVE, ids=str)
def test_active_dims(be, ad):
    rng = np.random.default_rng(11)
    x, y = rng.random((33, 5)), rng.random((34, 5))
    K = np.asarray(C.Matern32(0.9, active_dims=ad)(x, y))
    np.testing.assert_allclose(K, O.Matern32(0.9, active_dims=ad)(x, y), rtol=0, atol=1e-13)


def test_algebra_hierarchy(be):
    """Add / Mul / Pow with scalars and nested active_dims (tests/test_base_cov.py:120-160 shape)."""
    rng = np.random.default_rng(5)
    x, y = rng.random((40, 6)), rng.random((37, 6))
    k = (C.Matern52(1.2, active_dims=slice(None, -1)) * C.ExpQuad(0.4, active_dims=-1) + 0.3) ** 2
    ko = (O.Matern52(1.2, active_dims=slice(None, -1)) * O.ExpQuad(0.4, active_dims=-1) + 0.3) ** 2
    np.testing.assert_allclose(np.asarray(k(x, y)), ko(x, y), rtol=1e-13, atol=1e-13)
    k2 = 0.2 + C.Linear(2.0, active_dims=[0, 1]) * 1.5 + C.Exponential(0.8) * C.RatQuad(1.5, 0.9, active_dims=[2, 4])
    k2o = 0.2 + O.Linear(2.0, active_dims=[0, 1]) * 1.5 + O.Exponential(0.8) * O.RatQuad(1.5, 0.9, active_dims=[2, 4])
    np.testing.assert_allclose(np.asarray(k2(x, y)), k2o(x, y), rtol=1e-13, atol=1e-13)
    # outer active_dims, children index into the already-selected columns (base_cov.py:310-315)
    k3 = C.Matern32(0.7, active_dims=[0, 2]) + C.ExpQuad(1.1, active_dims=1)
    k3.active_dims = [1, 3, 5]
    k3o = O.Matern32(0.7, active_dims=[0, 2]) + O.ExpQuad(1.1, active_dims=1)
    k3o.active_dims = [1, 3, 5]
    np.testing.assert_allclose(np.asarray(k3(x, y)), k3o(x, y), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(k.diag(x), ko.diag(x), rtol=1e-13)


def test_covariances_are_mutable_after_first_use(be):
    """The reference's tests assign ``cov.active_dims`` (and users assign ``ls``) AFTER a first evaluation
    (tests/test_base_cov.py:31-37): the compiled device program must follow the object's current state."""
    rng = np.random.default_rng(6)
    x, y = rng.random((30, 5)), rng.random((21, 5))
    k, ko = C.Matern32(1.4) + C.Exponential(3.4), O.Matern32(1.4) + O.Exponential(3.4)
    np.testing.assert_allclose(np.asarray(k(x, y)), ko(x, y), rtol=1e-13)
    for ad in (slice(2), 1, slice(None, None, 2), [0, 3]):
        k.active_dims = ad
        ko.active_dims = ad
        np.testing.assert_allclose(np.asarray(k(x, y)), ko(x, y), rtol=1e-13)
    leaf, leafo = C.ExpQuad(0.9), O.ExpQuad(0.9)
    np.testing.assert_allclose(np.asarray(leaf(x, y)), leafo(x, y), rtol=1e-13)
    leaf.ls, leafo.ls = 2.5, 2.5
    np.testing.assert_allclose(np.asarray(leaf(x, y)), leafo(x, y), rtol=1e-13)
    back = C.Covariance.from_json(k.to_json())
    np.testing.assert_allclose(np.asarray(back(x, y)), ko(x, y), rtol=1e-13)


def test_empty_inputs(be):
    x = np.zeros((0, 3))
    y = np.random.default_rng(0).random((4, 3))
    assert np.asarray(C.Matern52(1.0)(x, y)).shape == (0, 4)
    assert np.asarray(C.Matern52(1.0)(y, x)).shape == (4, 0)


def _spd(n, seed, cond=1e6):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.geomspace(1.0, 1.0 / cond, n)
    return (q * ev) @ q.T


@pytest.mark.parametrize("variant", [0, 1], ids=["leaf128", "leaf32"])
@pytest.mark.parametrize("n", [1, 5, 32, 33, 100, 128, 129, 257, 1000, 1700])
def test_potrf(be, variant, n):
    A = _spd(n, n)
    Ad = be.upload(A.copy())
    be.set_option("trsm", variant)
    try:
        assert be.potrf(Ad) == 0
    finally:
        be.set_option("trsm", 0)
    Lc = Ad.numpy()
    Lref = np.linalg.cholesky(A)
    assert np.allclose(np.triu(Lc, 1), 0.0)
    assert rel_err(Lc @ Lc.T, A) < 1e-13
    assert rel_err(Lc, Lref) < 1e-9


def test_potrf_graph_replay(be):
    """n >= 1024 replays a captured CUDA graph on context-owned buffers: several matrices through the same graph,
    a second size, a non-positive pivot found inside a replay, and the stream path (graph off) for comparison."""
    for seed, n in ((1, 1300), (2, 1300), (3, 1100), (4, 1300)):
        A = _spd(n, seed)
        for graph in (1, 0):
            be.set_option("graph", graph)
            try:
                Ad = be.upload(A.copy())
                assert be.potrf(Ad) == 0
            finally:
                be.set_option("graph", 1)
            Lc = Ad.numpy()
            assert np.allclose(np.triu(Lc, 1), 0.0)
            assert rel_err(Lc @ Lc.T, A) < 1e-13
    A = _spd(1300, 5)
    A[700, 700] = -1.0
    info = be.potrf(be.upload(A))
    assert (641 if be.name == "cuda" else 1) <= info <= 701
    assert be.potrf(be.upload(_spd(1300, 6))) == 0   # the flag of the failed replay does not stick


def test_potrf_reports_non_positive_definite(be):
    A = _spd(64, 1)
    A[40, 40] = -1.0
    info = be.potrf(be.upload(A))
    assert 1 <= info <= 41
    A = _spd(300, 1)
    A[200, 200] = -1.0
    info = be.potrf(be.upload(A))
    # the CUDA factorisation reports the failing pivot of the second 128-wide leaf; the NumPy test double only flags failure
    assert (129 if be.name == "cuda" else 1) <= info <= 201


@pytest.mark.parametrize("variant", [0, 1], ids=["inv128-gemm-leaves", "substitution-leaves"])
@pytest.mark.parametrize("n,m", [(1, 1), (10, 33), (500, 100), (131, 257), (513, 1), (700, 128), (1000, 129),
                                 (2000, 300), (3000, 1000)])
def test_trsm_right(be, variant, n, m):
    rng = np.random.default_rng(n + m)
    Lp = np.linalg.cholesky(_spd(m, m, 1e4))
    X = rng.standard_normal((n, m))
    be.set_option("trsm", variant)
    try:
        out = be.trsm_right_lt(be.upload(Lp), be.upload(X.copy())).numpy()
    finally:
        be.set_option("trsm", 0)
    ref = solve_triangular(Lp, X.T, lower=True).T
    assert rel_err(out, ref) < 1e-11


@pytest.mark.parametrize("m,nrhs", [(1, 1), (40, 1), (100, 3), (257, 1), (300, 70), (97, 129), (384, 1), (700, 1), (1301, 1)])
@pytest.mark.parametrize("trans", [False, True])
def test_tri_solve(be, m, nrhs, trans):
    rng = np.random.default_rng(m * 7 + nrhs)
    Lp = np.linalg.cholesky(_spd(m, m + 1, 1e4))
    B = rng.standard_normal((m, nrhs)) if nrhs > 1 else rng.standard_normal(m)
    out = be.tri_solve(be.upload(Lp), B, trans=trans)
    ref = solve_triangular(Lp.T if trans else Lp, B, lower=not trans)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 1e-11


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (17, 9, 5), (128, 128, 16), (130, 257, 77), (300, 65, 1000)])
def test_gemm(be, ta, tb, m, n, k):
    rng = np.random.default_rng(m + n + k)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    out = be.gemm(A, B, trans_a=bool(ta), trans_b=bool(tb)).numpy()
    ref = (A.T if ta else A) @ (B.T if tb else B)
    assert rel_err(out, ref) < 1e-13


@pytest.mark.run_last
@pytest.mark.parametrize("m,k", [(5, 3), (130, 77), (257, 1000)])
def test_gemm_accumulates_into_its_output(be, m, k):
    """``alpha A A^T + beta C`` with the operand passed twice: the landmark leverage's ``sigma^2 Lp Lp^T + B^T B``."""
    rng = np.random.default_rng(m + k)
    A, C0 = rng.standard_normal((m, k)), rng.standard_normal((m, m))
    Ad = be.upload(A)
    out = be.gemm(Ad, Ad, trans_b=True, alpha=0.37, beta=1.0, out=be.upload(C0.copy())).numpy()
    assert rel_err(out, 0.37 * A @ A.T + C0) < 1e-13
    out = be.gemm(Ad, be.eye(k), alpha=-2.0, beta=0.5, out=be.upload(A.copy())).numpy()
    assert rel_err(out, -2.0 * A + 0.5 * A) < 1e-13


@pytest.mark.run_last
@pytest.mark.parametrize("n,c", [(1, 1), (7, 3), (300, 129), (1025, 64)])
def test_row_scaling_and_diagonal_vector(be, n, c):
    """``mb_mat_scale_rows`` / ``mb_mat_add_diag_vec``: the per-observation noise forms of the regression path."""
    rng = np.random.default_rng(n * 3 + c)
    A, s = rng.standard_normal((n, c)), rng.random(n) + 0.1
    assert np.array_equal(be.scale_rows(be.upload(A.copy(), sharded=True), s).numpy(), A * s[:, None])
    S, v = rng.standard_normal((n, n)), rng.random(n)
    assert np.array_equal(be.add_diag_vec(be.upload(S.copy()), v).numpy(), S + np.diag(v))


@pytest.mark.parametrize("variant", [0, 1, 2, 3], ids=["dmma16w", "dfma", "dmma8w", "nosplit"])
@pytest.mark.parametrize("n,r", [(1, 1), (50, 7), (1000, 64), (777, 130), (3000, 257), (20000, 1500)])
def test_gram_and_ridge(be, variant, n, r):
    if n >= 20000 and variant in (1, 2):
        pytest.skip("the large split-k case is for the default kernel and its no-split twin")
    rng = np.random.default_rng(n + r)
    L = rng.standard_normal((n, r)) / np.sqrt(r)
    t = rng.standard_normal(n)
    be.set_option("gemm", variant)
    try:
        Ld = be.upload(L, sharded=True)
        G = be.gram(Ld).numpy()
        assert rel_err(G, L.T @ L) < 1e-13
        assert np.array_equal(G, G.T)
        assert rel_err(be.gemv_t(Ld, t), L.T @ t) < 1e-13
        if r <= n:
            z0 = be.ridge_init(Ld, t)
            assert rel_err(z0, O.ridge_normal_equations(L, t)) < 1e-10
    finally:
        be.set_option("gemm", 0)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("n,r", [(1, 1), (9, 4), (3, 2), (1000, 64), (513, 33), (2000, 1000), (300, 2050), (4000, 5000),
                                 (5000, 2000), (777, 5000), (20011, 5000), (100, 8192), (64, 9000)])
def test_loss_grad_hess_transform(be, variant, n, r):
    if n >= 20000 and variant != 0:
        pytest.skip("the 800 MB case is for the default (bulk-TMA ring) kernel")
    rng = np.random.default_rng(n * 3 + r)
    L = rng.standard_normal((n, r)) / np.sqrt(r)
    nn = rng.random(n) * 0.5 + 0.05
    d, mu = 7.0, -3.0
    z = rng.standard_normal(r) * 0.3
    V, Vdr = O.nn_constants(nn, d)
    be.set_option("lossgrad", variant)
    try:
        st = be.objective(L, V, float(np.sum(Vdr)), mu, r)
        loss, grad = be.loss_grad(st, z)
        lref, gref = O.loss_and_grad(L, nn, d, mu, z, r)
        assert abs(loss - lref) <= 1e-12 * abs(lref)
        assert rel_err(grad, gref) < 1e-12
        assert rel_err(be.hess_diag(st, z), O.hessian_diag(L, nn, d, mu, z)) < 1e-12
        assert rel_err(be.transform(st.L, z, mu), L @ z + mu) < 1e-13
        # bit-reproducible run to run
        loss2, grad2 = be.loss_grad(st, z)
        assert loss2 == loss and np.array_equal(grad, grad2)
    finally:
        be.set_option("lossgrad", 0)


def test_loss_with_per_cell_d(be):
    rng = np.random.default_rng(9)
    n, r = 400, 50
    L = rng.standard_normal((n, r)) / np.sqrt(r)
    nn = rng.random(n) * 0.5 + 0.05
    dvec = rng.integers(2, 9, n).astype(float)
    z = rng.standard_normal(r) * 0.2
    V, Vdr = O.nn_constants(nn, dvec)
    st = be.objective(L, V, float(np.sum(Vdr)), -2.0, r)
    loss, grad = be.loss_grad(st, z)
    lref, gref = O.loss_and_grad(L, nn, dvec, -2.0, z, r)
    assert abs(loss - lref) <= 1e-12 * abs(lref)
    assert rel_err(grad, gref) < 1e-12


@pytest.mark.parametrize("nq,m,d,p", [(1, 1, 1, 1), (100, 37, 5, 1), (1000, 500, 50, 1), (300000, 64, 10, 1),
                                      (257, 129, 7, 3)])
def test_predict_mean(be, nq, m, d, p):
    rng = np.random.default_rng(nq + m)
    xq, xu = rng.random((nq, d)), rng.random((m, d))
    w = rng.standard_normal((m, p)) if p > 1 else rng.standard_normal(m)
    cov, covo = C.Matern52(0.8 * np.sqrt(d)), O.Matern52(0.8 * np.sqrt(d))
    ref = 1.5 + covo(xq, xu) @ w
    for variant in (0, 1):
        be.set_option("cov", variant)
        try:
            out = be.predict_mean(cov, xq, xu, w, 1.5)
        finally:
            be.set_option("cov", 0)
        assert out.shape == ref.shape
        assert rel_err(out, ref) < 1e-12


def test_predict_mean_product_kernel(be):
    rng = np.random.default_rng(4)
    xq, xu, w = rng.random((500, 6)), rng.random((90, 6)), rng.standard_normal(90)
    cov = C.Matern32(1.1, active_dims=slice(None, -1)) * C.ExpQuad(0.5, active_dims=-1)
    covo = O.Matern32(1.1, active_dims=slice(None, -1)) * O.ExpQuad(0.5, active_dims=-1)
    assert rel_err(be.predict_mean(cov, xq, xu, w, -0.5), -0.5 + covo(xq, xu) @ w) < 1e-12


@pytest.mark.parametrize("n", [1, 10, 200, 513])
def test_eigh(be, n):
    A = _spd(n, n + 3, 1e8)
    w, V = be.eigh(be.upload(A.copy()))
    V = V.numpy()
    wref = np.linalg.eigvalsh(A)
    assert np.max(np.abs(w - wref)) < 1e-13 * wref.max()
    assert rel_err((V * w) @ V.T, A) < 1e-12
    assert rel_err(V.T @ V, np.eye(n)) < 1e-12


def test_small_matrix_ops(be):
    rng = np.random.default_rng(2)
    A = rng.standard_normal((37, 21))
    Ad = be.upload(A)
    assert np.array_equal(be.transpose(Ad).numpy(), A.T)
    assert np.array_equal(be.copy_cols(Ad, 5, 9).numpy(), A[:, 5:14])
    s = rng.random(21)
    assert np.allclose(be.scale_cols(be.copy(Ad), s).numpy(), A * s, rtol=1e-15)
    assert np.allclose(be.scale(be.copy(Ad), 0.25).numpy(), A * 0.25, rtol=0, atol=0)
    assert rel_err(be.row_sumsq(Ad), np.sum(A * A, axis=1)) < 1e-14
    assert np.array_equal(be.eye(5).numpy(), np.eye(5))


@pytest.mark.parametrize("n,d", [(2, 1), (50, 3), (1000, 10), (3000, 50)])
def test_nn_distances_match_exact_search(be, n, d):
    """tests/test_parameters.py:244-268 semantics: exact nearest neighbour; index selection bit-exact."""
    from sklearn.neighbors import NearestNeighbors

    x = np.random.default_rng(n + d).random((n, d))
    dist, idx = be.nn_distances(x, return_index=True)
    rd, ri = NearestNeighbors(n_neighbors=2, algorithm="brute").fit(x).kneighbors(x)
    assert np.array_equal(idx, ri[:, 1])
    np.testing.assert_allclose(dist, rd[:, 1], rtol=1e-13)


def test_nn_distances_known_answers(be):
    """The reference's own known-answer cases (tests/test_parameters.py:244-268)."""
    x = np.array([[0.0, 0.0], [1.0, 1.0], [2.0, 2.0]])
    np.testing.assert_allclose(be.nn_distances(x), np.sqrt(2.0) * np.ones(3), rtol=1e-15)
    x = np.array([[1.0, 1.0], [1.0, 1.0], [1.0, 1.0]])
    np.testing.assert_array_equal(be.nn_distances(x), np.zeros(3))


# ---- expressions that do not fit ONE device program: split at the root, combined on the device -----------------
class _UserKernel(mb.cov.Covariance):
    """A user-defined kernel (subclasses overriding k must keep working)."""

    def k(self, x, y):
        return np.exp(-np.abs(np.asarray(x)[:, None, 0] - np.asarray(y)[None, :, 0]))


class _ExtendedMatern(mb.cov.Matern52):
    def k(self, x, y):
        return 2.0 * super().k(x, y)


def test_large_and_user_defined_expressions_are_split_not_evaluated_on_the_host(be):
    rng = np.random.default_rng(5)
    x, y = rng.random((70, 4)), rng.random((33, 4))
    leaves = [mb.cov.Matern52(1.0), mb.cov.Matern32(2.0), mb.cov.ExpQuad(0.5), mb.cov.Matern52(3.0, active_dims=[0, 2]),
              mb.cov.ExpQuad(1.5), mb.cov.Exponential(0.7)]
    six = leaves[0] + leaves[1] + leaves[2] * leaves[3] + leaves[4] * 0.3 + leaves[5] ** 2.0       # 6 leaves > MB_MAX_LEAVES
    oleaves = [O.Matern52(1.0), O.Matern32(2.0), O.ExpQuad(0.5), O.Matern52(3.0, active_dims=[0, 2]), O.ExpQuad(1.5),
               O.Exponential(0.7)]
    ref = (oleaves[0](x, y) + oleaves[1](x, y) + oleaves[2](x, y) * oleaves[3](x, y) + oleaves[4](x, y) * 0.3
           + oleaves[5](x, y) ** 2.0)
    assert not be.supports(six, 4)
    before = list(getattr(be.lib, "calls", []))
    np.testing.assert_allclose(six(x, y), ref, rtol=0, atol=5e-13)
    if hasattr(be.lib, "calls"):   # the test double records the entry points: every leaf went through mb_cov_build
        assert be.lib.calls[len(before):].count("mb_cov_build") >= 2
    refd = sum(float(np.squeeze(k(x[:1], x[:1]))) for k in oleaves[:2]) + float(np.squeeze(oleaves[2](x[:1], x[:1]) * oleaves[3](x[:1], x[:1]))) \
        + 0.3 * float(np.squeeze(oleaves[4](x[:1], x[:1]))) + float(np.squeeze(oleaves[5](x[:1], x[:1]))) ** 2
    np.testing.assert_allclose(six.diag(x)[0], refd, rtol=1e-12)
    # a user kernel inside Mul / Add / Pow, and a user class extending a stock kernel through super().k
    user = _UserKernel()
    np.testing.assert_allclose((user * mb.cov.Matern52(1.0))(x, y), user.k(x, y) * O.Matern52(1.0)(x, y), atol=5e-13)
    np.testing.assert_allclose((user + 1.0)(x, y), user.k(x, y) + 1.0, atol=5e-13)
    np.testing.assert_allclose(((user + mb.cov.Matern32(1.0)) ** 2.0)(x, y), (user.k(x, y) + O.Matern32(1.0)(x, y)) ** 2,
                               atol=5e-13)
    np.testing.assert_allclose(_ExtendedMatern(1.0)(x, y), 2.0 * O.Matern52(1.0)(x, y), atol=5e-13)
    np.testing.assert_allclose((user * mb.cov.Matern52(1.0)).diag(x), np.ones(70) * O.Matern52(1.0)(x[:1], x[:1])[0, 0],
                               rtol=1e-12)


def test_estimator_fits_with_an_expression_too_large_for_one_program(be):
    rng = np.random.default_rng(6)
    X = rng.random((400, 3))
    nn = O.compute_nn_distances(X)
    lm = X[:25].copy()
    cov = mb.cov.Matern52(0.9) + mb.cov.Matern32(1.1) + mb.cov.ExpQuad(0.8) + mb.cov.Matern52(2.0) + mb.cov.ExpQuad(1.7) * 0.5
    ocov = O.Add(O.Add(O.Add(O.Add(O.Matern52(0.9), O.Matern32(1.1)), O.ExpQuad(0.8)), O.Matern52(2.0)),
                 O.Mul(O.ExpQuad(1.7), 0.5))
    est = mb.DensityEstimator(cov_func=cov, landmarks=lm, nn_distances=nn)
    dens = est.fit_predict(X)
    ref = O.fit_density(X, cov_func=ocov, landmarks=lm, nn_distances=nn)
    np.testing.assert_allclose(dens, ref.log_density_x, rtol=1e-5)
    np.testing.assert_allclose(est.predict(X[:50]), O.predict_density(ref, X, X[:50]), rtol=1e-5)


# ---- K4 on tcgen05 kind::i8 digit slices (csrc/mb_i8.cu): chunks of >= 2048 cells, r >= 512 ---------------------
@pytest.mark.parametrize("n,r", [(65536, 512), (70001, 640), (66003, 777), (131072, 1030)])
def test_gram_on_int8_digit_slices(be, n, r):
    """The Gram matrix of a tall factor through the int8 digit-slice path: every entry within float64 rounding of the
    exact product, measured against (|L|^T |L|)_ij like a float64 dot product would be; columns of very different
    magnitude (a whitened covariance block) and ragged shapes; identical to rounding with the FP64 DMMA path."""
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only (the test double's Gram is NumPy's)")
    rng = np.random.default_rng(n + r)
    colscale = 10.0 ** (-6.0 * np.arange(r) / r)
    L = (rng.random((n, 1)) - 0.5 + 0.3 * (rng.random((n, r)) - 0.5)) * colscale
    Ld = be.upload(L, sharded=True)
    G = be.gram(Ld).numpy()
    be.set_option("i8", 0)
    try:
        G64 = be.gram(Ld).numpy()
    finally:
        be.set_option("i8", 1)
    ref = L.T @ L
    bound = np.abs(L).T @ np.abs(L)
    assert np.array_equal(G, G.T)
    # float64 accumulation over n terms (the reference itself) is good to ~ sqrt(n) eps of the bound
    assert np.max(np.abs(G - ref) / bound) < 2e-14
    assert np.max(np.abs(G - G64) / bound) < 2e-14
    t = rng.standard_normal(n)
    z0 = be.ridge_init(Ld, t)
    assert rel_err(z0, O.ridge_normal_equations(L, t)) < 1e-8


def test_int8_packs_on_the_side_stream_give_identical_bits(be):
    """The digit pack of the next slab (TRSM updates, > 65536 rows) / chunk (Gram) runs on a side stream under the MMA
    kernel of the current one with ``mb_set_option("i8_overlap", 1)``: same kernels on the same data, so the bits are
    those of the one-stream sequence (the default), run after run, and the solve is right."""
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only")
    rng = np.random.default_rng(11)
    n, m = 140003, 640                                    # three slabs of rows, updates with k = 384 and k = 256... on int8
    Lp = np.linalg.cholesky(_spd(m, m, 1e4))
    X = rng.standard_normal((n, m)) * 10.0 ** rng.uniform(-4, 2, size=(n, 1))
    Lpd = be.upload(Lp)
    outs, grams = [], []
    for overlap in (1, 0, 1):
        be.set_option("i8_overlap", overlap)
        try:
            Xd = be.upload(X.copy(), sharded=True)
            Ld = be.trsm_right_lt(Lpd, Xd)
            outs.append(Ld.numpy())
            grams.append(be.gram(Ld).numpy())
        finally:
            be.set_option("i8_overlap", 0)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert np.array_equal(grams[0], grams[1]) and np.array_equal(grams[0], grams[2])
    ref = solve_triangular(Lp, X.T, lower=True).T
    rownorm = np.max(np.abs(ref), axis=1, keepdims=True)
    assert np.max(np.abs(outs[0] - ref) / rownorm) < 1e-11
    bound = np.abs(outs[0]).T @ np.abs(outs[0])
    assert np.max(np.abs(grams[0] - outs[0].T @ outs[0]) / bound) < 2e-14


def test_gram_int8_rejects_non_finite_operands(be):
    L = np.random.default_rng(0).random((65536, 512))
    L[70, 3] = np.inf
    if be.name == "cuda":
        with pytest.raises(Exception, match="non-finite"):
            be.gram(be.upload(L, sharded=True))


@pytest.mark.parametrize("n,m", [(8192, 1024), (20011, 1500), (9000, 2600)])
def test_trsm_right_with_int8_updates(be, n, m):
    """K3 at sizes where the large off-diagonal updates X2 -= X1 L21^T run on the tcgen05 int8 digit slices
    (>= 8192 rows, k >= 512, >= 256 output columns): same tolerance as the FP64 path, and the two agree to rounding.
    Rows of very different magnitude (every row carries its own scale) and a ragged row count."""
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only")
    rng = np.random.default_rng(n + m)
    Lp = np.linalg.cholesky(_spd(m, m, 1e4))
    X = rng.standard_normal((n, m)) * 10.0 ** rng.uniform(-4, 2, size=(n, 1))
    Lpd = be.upload(Lp)
    out = be.trsm_right_lt(Lpd, be.upload(X.copy())).numpy()
    be.set_option("i8", 0)
    try:
        out64 = be.trsm_right_lt(Lpd, be.upload(X.copy())).numpy()
    finally:
        be.set_option("i8", 1)
    ref = solve_triangular(Lp, X.T, lower=True).T
    rownorm = np.max(np.abs(ref), axis=1, keepdims=True)
    assert np.max(np.abs(out - ref) / rownorm) < 1e-11
    assert np.max(np.abs(out - out64) / rownorm) < 1e-11
    # the sharded view of the same rows gives the same bits (the path is chosen from the global row count)
    half = be.trsm_right_lt(Lpd, be.upload(X[: n // 2].copy())).numpy()
    if n // 2 >= 8192:
        assert np.array_equal(half, out[: n // 2])


@pytest.mark.parametrize("n,k,p,trans_b", [(9000, 1100, 700, False), (8200, 640, 300, True)])
def test_tall_gemm_with_int8_slices(be, n, k, p, trans_b):
    """C = A B for a tall row-sharded A (cells x k): the Nystroem factor Q V (decomposition.py:265) on the int8
    digit slices; rows of very different magnitude, agreement with NumPy and with the FP64 DMMA path to rounding."""
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only")
    rng = np.random.default_rng(n + k + p)
    A = rng.standard_normal((n, k)) * 10.0 ** rng.uniform(-3, 3, size=(n, 1))
    B = rng.standard_normal((p, k) if trans_b else (k, p)) * 10.0 ** rng.uniform(-2, 2, size=(p, 1) if trans_b else (1, p))
    Ad = be.upload(A, sharded=True)
    out = be.gemm(Ad, be.upload(B), trans_b=trans_b).numpy()
    be.set_option("i8", 0)
    try:
        out64 = be.gemm(Ad, be.upload(B), trans_b=trans_b).numpy()
    finally:
        be.set_option("i8", 1)
    Bm = B.T if trans_b else B
    ref, bound = A @ Bm, np.abs(A) @ np.abs(Bm)
    assert np.max(np.abs(out - ref) / bound) < 1e-14
    assert np.max(np.abs(out - out64) / bound) < 1e-14


@pytest.mark.parametrize("n,r", [(1, 1), (50, 7), (777, 130), (3000, 257), (4099, 64)])
def test_int8_slices_forced_on_small_and_ragged_shapes(be, n, r):
    """Option i8 = 2 sends EVERY Gram / TRSM update / tall GEMM through the int8 digit-slice kernels: ragged and tiny
    shapes (partial tiles, partial k-steps, single rows) against NumPy at the FP64 tolerances."""
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only")
    rng = np.random.default_rng(n + r)
    L = rng.standard_normal((n, r)) / np.sqrt(r)
    be.set_option("i8", 2)
    try:
        Ld = be.upload(L, sharded=True)
        G = be.gram(Ld).numpy()
        assert rel_err(G, L.T @ L) < 1e-13 and np.array_equal(G, G.T)
        Lp = np.linalg.cholesky(_spd(r, r, 1e4))
        X = rng.standard_normal((n, r))
        out = be.trsm_right_lt(be.upload(Lp), be.upload(X.copy(), sharded=True)).numpy()
        assert rel_err(out, solve_triangular(Lp, X.T, lower=True).T) < 1e-11
        B = rng.standard_normal((r, 37))
        assert rel_err(be.gemm(Ld, be.upload(B)).numpy(), L @ B) < 1e-13
    finally:
        be.set_option("i8", 1)


@pytest.mark.parametrize("kc,ko", [(C.Matern52, O.Matern52), (C.Matern32, O.Matern32), (C.ExpQuad, O.ExpQuad),
                                   (C.Exponential, O.Exponential)])
@pytest.mark.parametrize("n,m,d", [(300, 70, 10), (129, 65, 64), (1000, 257, 33), (5000, 300, 50), (64, 64, 1)])
def test_k1_on_int8_digit_slices(be, kc, ko, n, m, d):
    """K1 with its contraction on the tcgen05 int8 digit slices (csrc/mb_cov_i8.cu), forced at small and ragged shapes
    (option cov_i8 = 2): partial cell panels, partial landmark tiles, one and two k-steps, all four kernels; the same
    2e-13 bar as the FP64 kernels, on unit-scale data where the length scale does not hide the contraction."""
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only")
    rng = np.random.default_rng(n + m + d)
    x, y = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    if kc is not C.Exponential:
        # coincident points: d = sqrt(1e-12).  Not for the Exponential kernel: exp(-d / (2 ls)) is not smooth at d = 0,
        # so there the rounding of xx - 2xy + yy (1e-16 of |x|^2, in ANY float64 implementation, the reference's
        # included) shows up as 1e-9 in k, whatever computes it
        y[: min(m, 5)] = x[: min(m, 5)]
    be.set_option("cov_i8", 2)
    try:
        for ls in (1.3, 25.0):
            K = be.cov(kc(ls), x, y).numpy()
            assert np.max(np.abs(K - ko(ls)(x, y))) < 2e-13
        sel = [0] if d == 1 else list(range(0, d, 2))
        K = be.cov(kc(2.0, active_dims=sel), x, y).numpy()
        assert np.max(np.abs(K - ko(2.0, active_dims=sel)(x, y))) < 2e-13
    finally:
        be.set_option("cov_i8", 1)


def test_k1_int8_large_shape_matches_the_dmma_kernel(be):
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only")
    rng = np.random.default_rng(11)
    x, y = rng.random((20011, 50)), rng.random((1003, 50))
    K8 = be.cov(C.Matern52(38.0), x, y).numpy()
    be.set_option("cov_i8", 0)
    try:
        K64 = be.cov(C.Matern52(38.0), x, y).numpy()
    finally:
        be.set_option("cov_i8", 1)
    assert np.max(np.abs(K8 - K64)) < 2e-13
    assert np.max(np.abs(K8 - O.Matern52(38.0)(x, y))) < 2e-13


def test_int8_gemm_issue_variants_give_identical_bits(be):
    """The int8 digit-slice GEMMs with 1, 2 or 4 MMA-issuing warps and with the A operand staged in tensor memory
    (option i8_issuers = 0): the integer accumulation is exact, so every variant must return the same bits."""
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only")
    rng = np.random.default_rng(3)
    L = rng.standard_normal((66000, 640)) * 10.0 ** rng.uniform(-3, 1, size=(1, 640))
    Lp = np.linalg.cholesky(_spd(640, 640, 1e4))
    Ld, Lpd = be.upload(L, sharded=True), be.upload(Lp)
    ref_g = ref_t = None
    try:
        for iss in (4, 2, 1, 0):
            be.set_option("i8_issuers", iss)
            G = be.gram(Ld).numpy()
            T = be.trsm_right_lt(Lpd, be.copy(Ld)).numpy()
            if ref_g is None:
                ref_g, ref_t = G, T
                assert rel_err(G, L.T @ L) < 1e-13
                assert rel_err(T, solve_triangular(Lp, L.T, lower=True).T) < 1e-11
            else:
                assert np.array_equal(G, ref_g), f"Gram differs with i8_issuers={iss}"
                assert np.array_equal(T, ref_t), f"TRSM differs with i8_issuers={iss}"
    finally:
        be.set_option("i8_issuers", 4)


@pytest.mark.parametrize("n,d,forced", [(2, 1, True), (50, 3, True), (1000, 10, True), (3000, 50, True), (777, 64, True),
                                        (20011, 50, False), (9000, 33, False)])
def test_nn_distances_on_int8_digit_slices(be, n, d, forced):
    """Exact nearest neighbour through the tcgen05 int8 digit-slice kernel (running-minimum epilogue): indices bit-exact
    against scikit-learn's brute-force search, duplicates and ragged shapes included; forced at small sizes."""
    if be.name != "cuda":
        pytest.skip("the int8 digit-slice path exists in the CUDA library only")
    from sklearn.neighbors import NearestNeighbors

    x = np.random.default_rng(n + d).random((n, d))
    if n >= 50:
        x[7] = x[3]                                           # a duplicate: distance 0, lowest index wins
    be.set_option("cov_i8", 2 if forced else 1)
    try:
        dist, idx = be.nn_distances(x, return_index=True)
    finally:
        be.set_option("cov_i8", 1)
    rd, ri = NearestNeighbors(n_neighbors=2, algorithm="brute").fit(x).kneighbors(x)
    exact = np.sqrt(np.sum((x - x[idx]) ** 2, axis=1))
    np.testing.assert_allclose(dist, exact, rtol=1e-13, atol=1e-300)
    # scikit-learn's brute force forms xx - 2xy + yy: a duplicated point comes out at ~1e-7 there, at exactly 0 here
    np.testing.assert_allclose(dist, rd[:, 1], rtol=1e-13, atol=1e-6)
    same = idx == ri[:, 1]
    # scikit-learn breaks exact ties (duplicated points) by its own order: accept any index at the same distance
    assert np.all(same | (np.abs(exact - rd[:, 1]) <= 1e-6))
    # rows whose nearest neighbour IS the duplicated point see a two-way tie as well: a handful of rows at most
    assert np.count_nonzero(~same) <= 8
