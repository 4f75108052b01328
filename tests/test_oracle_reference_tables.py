"""Pin the CPU oracle (and the host-side functions of the package that share those semantics)
against every known-answer table the reference's own tests hold for this path:

  tests/test_parameters.py:244-268  compute_nn_distances known answers / error cases
  tests/test_parameters.py:271-290  compute_gp_type truth table
  tests/test_parameters.py:318-364  compute_rank / compute_n_landmarks truth tables
  tests/test_util.py:59-89          test_rank == 4 on a constructed spectrum
  tests/test_laplace.py:40-82       Laplace std on quadratic / quartic / flat losses
  tests/test_cov.py:21-64           output shapes for 6 kernels x 5 active_dims forms
  mellon/decomposition.py:57-75     the searchsorted rank rule (SURVEY.md Appendix A worked example)
"""

from enum import Enum

import numpy as np
import pytest

import mellon_b200 as mb
from mellon_b200.util import GaussianProcessType as G
from oracle import mellon_oracle as O


@pytest.mark.parametrize("impl", ["oracle", "package"])
def test_compute_nn_distances_known_answers(impl, be):
    if be.name != "fake":
        pytest.skip("host logic")
    f = O.compute_nn_distances if impl == "oracle" else mb.parameters.compute_nn_distances
    x = np.array([[1, 2], [2, 3], [3, 4]], float)
    assert np.allclose(f(x), np.sqrt(2) * np.ones(3))
    x = np.array([[1, 2], [1, 2], [1, 2]], float)
    assert np.allclose(f(x), np.zeros(3))
    x = np.array([[1, 1], [2, 2], [4, 4], [5, 5]], float)
    assert np.allclose(f(x), np.sqrt(2) * np.ones(4))
    with pytest.raises(ValueError):
        f(np.array([[1, 2]], float))
    with pytest.raises(ValueError):
        f(np.array([]))


GP_TABLE = [
    ((0, 100, 100), G.FULL), ((100, 1.0, 100), G.FULL), ((100, None, 100), G.FULL), ((100, 0, 100), G.FULL),
    ((100, 50, 100), G.FULL_NYSTROEM), ((100, 0.5, 100), G.FULL_NYSTROEM),
    ((50, 50, 100), G.SPARSE_CHOLESKY), ((50, 1.0, 100), G.SPARSE_CHOLESKY), ((50, None, 100), G.SPARSE_CHOLESKY),
    ((50, 0, 100), G.SPARSE_CHOLESKY), ((50, 25, 100), G.SPARSE_NYSTROEM), ((50, 0.5, 100), G.SPARSE_NYSTROEM),
]


@pytest.mark.parametrize("args,expected", GP_TABLE)
def test_compute_gp_type_truth_table(args, expected):
    assert mb.parameters.compute_gp_type(*args) == expected
    assert O.compute_gp_type(*args).value == expected.value


def test_compute_rank_and_n_landmarks_tables():
    for mod, T in ((mb.parameters, G), (O, O.GaussianProcessType)):
        assert mod.compute_rank(T.FULL_NYSTROEM) == 0.99
        assert mod.compute_rank(T.SPARSE_CHOLESKY) == 1.0
        assert mod.compute_rank(None) == 1.0
        assert mod.compute_n_landmarks(None, 100, np.ones((50, 2))) == 50
        assert mod.compute_n_landmarks(None, 100, None) == 100
        assert mod.compute_n_landmarks(T.FULL, 100, None) == 100
        assert mod.compute_n_landmarks(T.FULL_NYSTROEM, 100, None) == 100
        for n in (100, 80):
            assert mod.compute_n_landmarks(T.SPARSE_CHOLESKY, n, None) == 5000
            assert mod.compute_n_landmarks(T.SPARSE_NYSTROEM, n, None) == 5000

        class UnknownType(Enum):
            UNKNOWN = "unknown"

        assert mod.compute_n_landmarks(UnknownType.UNKNOWN, 100, None) == 100


def test_mu_and_ls_are_finite_floats():
    for mod in (mb.parameters, O):
        mu = mod.compute_mu(np.arange(1, 101, dtype=float), 4)
        ls = mod.compute_ls(np.arange(1, 100, dtype=float))
        assert isinstance(mu, float) and np.isfinite(mu)
        assert isinstance(ls, float) and np.isfinite(ls)
    nn = np.random.default_rng(0).random(1000) + 0.1
    assert mb.parameters.compute_mu(nn, 7) == O.compute_mu(nn, 7)
    assert mb.parameters.compute_ls(nn) == O.compute_ls(nn)
    assert O.compute_ls(nn) == pytest.approx(np.exp(np.mean(np.log(nn)) + 3))


def _spectrum_matrix():
    rng = np.random.default_rng(423)
    sv = np.array([3.0, 2.0, 1.5, 1.0, 0.4])
    U, _ = np.linalg.qr(rng.standard_normal((5, 5)))
    V, _ = np.linalg.qr(rng.standard_normal((10, 10)))
    S = np.zeros((5, 10))
    S[:5, :5] = np.diag(sv)
    return U @ S @ V.T


def test_test_rank_known_answer(be, capsys):
    """tests/test_util.py:59-89: singular values {3, 2, 1.5, 1, 0.4}, tol 0.5 -> rank 4."""
    L = _spectrum_matrix()
    assert O.test_rank(L, tol=0.5) == 4
    assert mb.util.test_rank(L, tol=0.5) == 4
    assert mb.util.test_rank(L, tol=1.2, threshold=0.5) == 3
    assert "approx. rank fraction" in capsys.readouterr().out
    with pytest.raises(ValueError):
        mb.util.test_rank(np.ones(3))
    with pytest.raises(TypeError):
        mb.util.test_rank(3.0)


def test_rank_selection_rule_worked_example():
    """decomposition.py:57-75: spectrum {6.5, 2, 1, 0.5}, rank 0.99 keeps 3 (strictly below target)."""
    from mellon_b200.decomposition import _select_rank

    s = np.array([0.5, 1.0, 2.0, 6.5])  # ascending, as eigh returns
    A = np.diag(s)
    for rank, keep in ((0.99, 3), (0.5, 1), (0.66, 1), (0.9, 2), (0.0001, 1), (2, 2), (10, 4)):
        assert _select_rank(s, rank) == keep
        sv, v = O._eigendecomposition(A, rank=rank)
        assert sv.shape[0] == keep == v.shape[1]
    # non-positive eigenvalues are dropped before counting
    assert _select_rank(np.array([-1.0, 0.0, 1.0, 3.0]), 10) == 2


LAPLACE_CASES = [
    (lambda z, p=np.array([1.0, 4.0, 9.0, 0.25]): 0.5 * np.sum(p * (z - np.array([1.0, -2.0, 0.5, 3.0])) ** 2),
     np.array([1.0, -2.0, 0.5, 3.0]), 1.0 / np.sqrt(np.array([1.0, 4.0, 9.0, 0.25]))),
    (lambda z: np.sum(z ** 4 + z ** 2), np.zeros(5), np.ones(5) / np.sqrt(2.0)),
]


@pytest.mark.parametrize("loss,z,expected", LAPLACE_CASES)
def test_laplace_std_known_answers(loss, z, expected):
    """tests/test_laplace.py:40-73 (quadratic: 1/sqrt(precision); z^4 + z^2 at 0: 1/sqrt(2))."""
    assert np.allclose(O.compute_laplace_std_numeric(loss, z), expected, atol=1e-5)
    assert np.allclose(mb.inference.compute_laplace_std(loss, z), expected, atol=1e-5)


def test_laplace_clipping_near_zero_curvature():
    """tests/test_laplace.py:75-82: flat loss -> clipped at 1e-8, finite."""
    flat = lambda z: 0.0 * np.sum(z)
    for std in (O.compute_laplace_std_numeric(flat, np.zeros(3)), mb.inference.compute_laplace_std(flat, np.zeros(3))):
        assert np.all(np.isfinite(std)) and np.allclose(std, 1e4)


def test_laplace_closed_form_equals_hessian_vector_route():
    """The closed form 1 + sum_i A_i L_ij^2 equals the diagonal the reference extracts with r
    Hessian-vector products (inference.py:311-317), on a dense small case."""
    rng = np.random.default_rng(1)
    L = rng.standard_normal((300, 12)) / 3
    nn = rng.random(300) + 0.1
    z = rng.standard_normal(12) * 0.2
    a = O.hessian_diag(L, nn, 5.0, -2.0, z)
    b = O.hessian_diag_dense(L, nn, 5.0, -2.0, z)
    np.testing.assert_allclose(a, b, rtol=1e-12)
    # and both equal second differences of the oracle's loss function
    loss = O.compute_loss_func(nn, 5.0, O.compute_transform(-2.0, L), 12)
    num = 1.0 / O.compute_laplace_std_numeric(loss, z, h=1e-3) ** 2
    np.testing.assert_allclose(a, num, rtol=1e-5)


def test_oracle_gradient_is_the_derivative_of_the_reference_loss():
    """loss_and_grad's analytic gradient (z + L^T(A - 1)) against central differences of the
    restated loss function of inference.py:189-190."""
    rng = np.random.default_rng(2)
    L = rng.standard_normal((200, 9)) / 3
    nn = rng.random(200) + 0.1
    z = rng.standard_normal(9) * 0.3
    loss = O.compute_loss_func(nn, 4.0, O.compute_transform(-1.0, L), 9)
    val, grad = O.loss_and_grad(L, nn, 4.0, -1.0, z, 9)
    assert val == pytest.approx(loss(z), rel=1e-14)
    num = np.array([(loss(z + h) - loss(z - h)) / 2e-6 for h in 1e-6 * np.eye(9)])
    np.testing.assert_allclose(grad, num, rtol=1e-6)


COVS = ["Matern32", "Matern52", "ExpQuad", "Exponential", "RatQuad", "Linear"]
ACTIVE_DIMS = [None, slice(2), 1, slice(None, None, 2), [1, 2]]


@pytest.mark.parametrize("name", COVS)
@pytest.mark.parametrize("active_dims", ACTIVE_DIMS, ids=str)
def test_covariance_shapes_and_values(name, active_dims, be):
    """tests/test_cov.py:21-64: (n, n + 1) output for every kernel x active_dims form; the package's
    result equals the oracle's."""
    n, d, ls = 5, 4, 1.2
    mk = lambda mod: getattr(mod, name)(3, ls, active_dims=active_dims) if name == "RatQuad" \
        else getattr(mod, name)(ls, active_dims=active_dims)
    cov, ocov = mk(mb.cov), mk(O)
    assert len(str(cov)) > 0
    x = np.ones((n, d))
    y = np.ones((n + 1, d)) * 2
    y[1] = 1.5
    values = np.asarray(cov(x, y))
    assert values.shape == (n, n + 1)
    np.testing.assert_allclose(values, ocov(x, y), rtol=1e-13, atol=1e-15)
    # k_grad against central differences of the oracle kernel (the reference checks against jacfwd)
    g = cov.k_grad(x)(y)
    assert g.shape == (n, n + 1, d)
    num = np.empty_like(g)
    for j in range(d):
        e = np.zeros(d)
        e[j] = 1e-6
        num[..., j] = (ocov(x, y + e) - ocov(x, y - e)) / 2e-6
    np.testing.assert_allclose(g, num, atol=1e-6)


def test_exponential_and_ratquad_conventions():
    """Appendix A: Exponential is exp(-r/2); RatQuad takes alpha first and uses exponent -alpha."""
    x, y = np.zeros((1, 1)), np.array([[2.0]])
    d = np.sqrt(4.0 + 1e-12)
    assert O.Exponential(1.0)(x, y)[0, 0] == pytest.approx(np.exp(-d / 2), rel=1e-15)
    assert O.RatQuad(3.0, 2.0)(x, y)[0, 0] == pytest.approx((d * d / 4 / 6 + 1) ** -3.0, rel=1e-14)
    assert O.Matern52(1.0)(x, x)[0, 0] < 1.0  # distance(x, x) = 1e-6, not 0
    assert O.distance(x, x)[0, 0] == pytest.approx(1e-6)


def test_fast_quantile_is_numpy_quantile_bit_for_bit():
    """compute_mu's percentile (parameters.py:586-599) comes from one selection instead of np.quantile's three: same bits."""
    from mellon_b200.parameters import _quantile_linear, compute_mu

    rng = np.random.default_rng(11)
    for trial in range(40):
        a = rng.standard_normal(int(rng.integers(4096, 60000))) * 10 ** rng.uniform(-3, 3)
        if trial % 5 == 0:
            a = np.round(a, 1)                                   # ties around the order statistic
        for q in (0.01, 0.5, 0.37, 0.999, 0.0, 1.0):
            assert _quantile_linear(a, q) == np.quantile(a, q)
    a[5] = np.nan
    assert np.isnan(_quantile_linear(a, 0.01))
    nn = rng.random(50000) * 0.3 + 0.01
    assert compute_mu(nn, 7) == O.compute_mu(nn, 7)
