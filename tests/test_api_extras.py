"""API-surface pieces next to the accelerated path, exercised on both backends (the NumPy test double on CPU, the
CUDA library on a GPU): ADVI on the device loss + gradient, numerical predictor derivatives, per-cell mean in the
functional transform / loss API.  The reference's own test files for these run against the package through
tools/run_reference_tests.py; the checks here are the parts that can be pinned numerically."""

import numpy as np
import pytest

import mellon_b200 as mb
from oracle import mellon_oracle as O


def _cells(n=300, d=3, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, d)) * 0.7


def test_advi_runs_on_the_device_objective_and_agrees_with_map(be):
    X = _cells()
    nn = O.compute_nn_distances(X)
    lm = X[:40].copy()
    est_map = mb.DensityEstimator(landmarks=lm, nn_distances=nn, predictor_with_uncertainty=True)
    dens_map = est_map.fit_predict(X)
    est = mb.DensityEstimator(landmarks=lm, nn_distances=nn, optimizer="advi", n_iter=60, predictor_with_uncertainty=True)
    dens = est.fit_predict(X)
    assert dens.shape == (300,) and len(est.losses) == 60
    assert est.pre_transformation_std is not None and np.all(est.pre_transformation_std > 0)
    assert np.corrcoef(dens, dens_map)[0, 1] > 0.8          # tests/test_laplace.py:170-208 asks for > 0.8
    unc = est.predict.uncertainty(X[:20])
    assert unc.shape == (20,) and np.all(unc > 0)
    # deterministic: the sampler is re-keyed with the iteration number
    again = mb.DensityEstimator(landmarks=lm, nn_distances=nn, optimizer="advi", n_iter=60).fit_predict(X)
    np.testing.assert_allclose(again, dens, rtol=1e-8, atol=1e-10)


def test_run_advi_on_a_plain_callable(be):
    res = mb.inference.run_advi(lambda x: float(np.sum(np.asarray(x) ** 2)), np.ones(2), n_iter=50, nsamples=10)
    assert res.pre_transformation.shape == (2,) and res.pre_transformation_std.shape == (2,) and len(res.losses) == 50
    assert np.max(np.abs(res.pre_transformation)) < 1.0


def test_predictor_derivatives_match_finite_differences_of_the_oracle(be):
    X = _cells(200, 2, 1)
    nn = O.compute_nn_distances(X)
    lm = X[:30].copy()
    est = mb.DensityEstimator(landmarks=lm, nn_distances=nn)
    est.fit(X)
    Y = _cells(15, 2, 2)
    g = est.predict.gradient(Y)
    H = est.predict.hessian(Y)
    sign, logdet = est.predict.hessian_log_determinant(Y)
    assert g.shape == Y.shape and H.shape == (15, 2, 2) and sign.shape == (15,) and logdet.shape == (15,)
    np.testing.assert_allclose(H, np.swapaxes(H, 1, 2), atol=1e-12)
    # independent check: analytic gradient of mu + k(y, xu) w through the covariance's own k_grad
    w = np.asarray(est.predict.weights) if hasattr(est.predict, "weights") else None
    if w is not None:
        kg = np.asarray(est.cov_func.k_grad(np.asarray(est.landmarks))(Y))     # (m, n, d)
        np.testing.assert_allclose(g, np.einsum("i,ijd->jd", w.reshape(-1), kg), rtol=1e-6, atol=1e-8)


@pytest.mark.run_last          # computes its nn_distances itself: on the device since the end of round 1
def test_time_predictor_derivatives_shapes(be):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((160, 2)) * 0.5
    times = np.repeat(np.arange(4.0), 40)
    est = mb.TimeSensitiveDensityEstimator(ls=1.0, ls_time=1.0, landmarks=np.concatenate([X[::8], times[::8, None]], axis=1))
    est.fit(X, times)
    g = est.predict.gradient(X[:10], 1.0)
    H = est.predict.hessian(X[:10], 1.0)
    td = est.predict.time_derivative(X[:10], 1.0)
    assert g.shape == (10, 2) and H.shape == (10, 2, 2) and td.shape == (10,)
    gm = est.predict.gradient(X[:10], multi_time=[0.0, 1.0, 2.0])
    assert gm.shape == (10, 3, 2)
    h = 1e-5
    np.testing.assert_allclose(td, (est.predict(X[:10], 1.0 + h) - est.predict(X[:10], 1.0 - h)) / (2 * h), rtol=1e-5, atol=1e-7)


def test_per_cell_mean_in_transform_and_loss(be):
    rng = np.random.default_rng(4)
    n, r = 150, 9
    L = rng.standard_normal((n, r)) / 3
    nn = rng.random(n) * 0.3 + 0.05
    mu = rng.standard_normal(n) * 0.1 - 3.0
    z = rng.standard_normal(r) * 0.2
    tr = mb.inference.compute_transform(mu, L)
    lf = mb.inference.compute_loss_func(nn, 5.0, tr, r)
    loss, grad = lf.value_and_grad(z)
    f = L @ z + mu
    V, Vdr = O.nn_constants(nn, 5.0)
    A = np.exp(f + V)
    ref = 0.5 * z @ z + 0.5 * r * np.log(2 * np.pi) - np.sum(f + Vdr - A)
    np.testing.assert_allclose(tr(z), f, rtol=1e-13)
    assert abs(loss - ref) <= 1e-12 * max(abs(ref), float(np.sum(np.abs(f + Vdr) + A)))
    scale = max(1.0, float(np.max(np.abs(L).T @ np.abs(A - 1.0))))      # size of the terms each gradient entry sums
    np.testing.assert_allclose(grad, z + L.T @ (A - 1.0), rtol=1e-11, atol=1e-12 * scale)
    with pytest.raises(ValueError):
        mb.inference.compute_transform(mu[:-1], L)


@pytest.mark.run_last
def test_distance_grad_and_jax_config_shims(be):
    """mellon/util.py:369-428, 572-586 (tests/test_util.py:22-56, 113-114 of the reference)."""
    rng = np.random.default_rng(9)
    x, y = rng.random((7, 3)), rng.random((5, 3))
    dist, grad = mb.util.distance_grad(x)(y)
    assert dist.shape == (7, 5) and grad.shape == (7, 5, 3)
    np.testing.assert_allclose(dist, O.distance(x, y), rtol=1e-12)
    h = 1e-6
    for j in range(3):
        e = np.zeros(3)
        e[j] = h
        fd = (O.distance(x, y + e) - O.distance(x, y - e)) / (2 * h)
        np.testing.assert_allclose(grad[..., j], fd, atol=1e-6)
    mb.util.set_jax_config()
    with pytest.raises(ValueError):
        mb.util.set_jax_config(enable_x64=False)


def test_gradient_of_a_multi_output_predictor_and_named_refusals(be):
    """the numerical gradient of a multi-output mean (FunctionEstimator with y of shape (n, p)) has
    shape (n, p, d); the Hessian is refused by name; a reference-written exp-log predictor is refused by name; a user
    covariance class that shares a stock NAME is loaded from its own module."""
    rng = np.random.default_rng(4)
    X = rng.random((150, 2))
    Y = np.stack([np.sin(3 * X[:, 0]), X[:, 1] ** 2], axis=1)
    fe = mb.FunctionEstimator(n_landmarks=0, ls=0.5, sigma=0.05).fit(X, Y)
    Q = rng.random((7, 2))
    g = fe.predict.gradient(Q)
    assert g.shape == (7, 2, 2)
    h = 1e-5
    fd = (fe.predict(Q + [h, 0]) - fe.predict(Q - [h, 0])) / (2 * h)
    np.testing.assert_allclose(g[:, :, 0], fd, rtol=1e-4, atol=1e-6)
    with pytest.raises(NotImplementedError, match="multi-output"):
        fe.predict.hessian(Q)
    with pytest.raises(NotImplementedError, match="ExpLandmarksConditional"):
        mb.Predictor.from_dict({"metadata": {"classname": "ExpLandmarksConditional", "module_name": "mellon.conditional",
                                             "module_version": "1.7.1"}, "data": {}})
