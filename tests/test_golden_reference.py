"""Golden vectors minted by the UNMODIFIED reference (oracle/make_golden.py runs /root/reference's own
source on the NumPy stand-ins of oracle/refshim; the reference's own golden-vector tests,
tests/test_reference_results.py, pass on that stack).

  * `test_oracle_*`   : the CPU oracle reproduces them  -> the oracle is pinned to the reference.
  * `test_package_*`  : the package reproduces them through the backend under test
                        (`fake` = host logic on CPU; `cuda`, marked gpu = the CUDA path on a B200).

Tolerances: kernels / decompositions / objective at fixed inputs are compared at rounding level;
end-to-end log densities at the reference's default L-BFGS-B stop within north_star's 1e-5 on the
benign (uniform) inputs and within the measured optimiser noise floor on clustered inputs, and at
1e-6 when both optimisers run to convergence (`*_tight`)."""

import os

import numpy as np
import pytest

import mellon_b200 as mb
from mellon_b200 import cov as C
from oracle import mellon_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TIGHT = dict(maxiter=20000, maxfun=100000, ftol=0.0, gtol=1e-9)


def load(name):
    return np.load(os.path.join(GOLDEN, f"reference_{name}.npz"))


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.abs(b)))


def kernel_objects(mod):
    k = {}
    for name in ("Matern32", "Matern52", "ExpQuad", "Exponential", "Linear"):
        k[name] = getattr(mod, name)(1.3)
        k[name + "_dims"] = getattr(mod, name)(0.7, active_dims=[0, 3])
    k["RatQuad"] = mod.RatQuad(2.5, 1.3)
    k["expr"] = (mod.Matern52(1.2, active_dims=slice(None, -1)) * mod.ExpQuad(0.4, active_dims=-1) + 0.3) ** 2
    nested = mod.Matern32(0.7, active_dims=[0, 2]) + mod.ExpQuad(1.1, active_dims=1)
    nested.active_dims = [1, 3, 4]
    k["nested"] = nested
    return k


# ---- kernels ------------------------------------------------------------------------------------------
def test_oracle_kernels_match_reference():
    g = load("kernels")
    for name, k in kernel_objects(O).items():
        np.testing.assert_allclose(k(g["x"], g["y"]), g[name], rtol=1e-14, atol=1e-15, err_msg=name)
    np.testing.assert_allclose(O.distance(g["x"], g["y"]), g["distance"], rtol=1e-14)
    np.testing.assert_allclose(O.Matern52(1.3).diag(g["x"]), g["diag_Matern52"], rtol=1e-15)


def test_package_kernels_match_reference(be):
    g = load("kernels")
    for name, k in kernel_objects(C).items():
        np.testing.assert_allclose(np.asarray(k(g["x"], g["y"])), g[name], rtol=2e-13, atol=1e-14, err_msg=name)
    np.testing.assert_allclose(np.asarray(mb.util.distance(g["x"], g["y"])), g["distance"], rtol=1e-13)
    np.testing.assert_allclose(C.Matern52(1.3).diag(g["x"]), g["diag_Matern52"], rtol=1e-14)


# ---- decompositions and the objective at fixed inputs ------------------------------------------------------
def test_oracle_stages_match_reference():
    g = load("stages")
    X, lm, nn = g["X"], g["landmarks"], g["nn_distances"]
    cov = O.Matern52(float(g["ls"]))
    np.testing.assert_allclose(O.compute_nn_distances(X), nn, rtol=1e-13)
    Lp = O.compute_Lp(X, cov, landmarks=lm)
    np.testing.assert_allclose(Lp, g["Lp"], rtol=1e-7, atol=1e-10)
    L = O.compute_L(X, cov, landmarks=lm, Lp=g["Lp"])
    np.testing.assert_allclose(L, g["L"], rtol=1e-7, atol=1e-9)
    for tag, rank, gp in (("nystroem15", 15, None), ("nystroem99", 0.99, O.GaussianProcessType.SPARSE_NYSTROEM)):
        Ln = O.compute_L(X, cov, landmarks=lm, rank=rank, gp_type=gp)
        assert list(Ln.shape) == list(g[f"L_{tag}_shape"].astype(int))          # rank selection bit-exact
        np.testing.assert_allclose(Ln[:60] @ Ln.T[:, ::5], g[f"L_{tag}_gram"], rtol=1e-6, atol=1e-8)
    Lf = O.compute_L(X[:120], cov, rank=0.9, gp_type=O.GaussianProcessType.FULL_NYSTROEM)
    assert list(Lf.shape) == list(g["L_full_nystroem_shape"].astype(int))
    np.testing.assert_allclose(Lf[:60] @ Lf.T[:, ::2], g["L_full_nystroem_gram"], rtol=1e-7, atol=1e-9)
    assert O.compute_mu(nn, 4) == pytest.approx(float(g["mu"]), rel=1e-14)
    assert O.compute_ls(nn) == pytest.approx(float(g["ls_heuristic"]), rel=1e-14)
    np.testing.assert_allclose(O.compute_initial_value(nn, 4, float(g["mu"]), g["L"]), g["initial_value"], rtol=1e-8)
    loss, grad = O.loss_and_grad(g["L"], nn, 4, float(g["mu"]), g["z"])
    assert loss == pytest.approx(float(g["loss"]), rel=1e-14)
    np.testing.assert_allclose(grad, g["grad"], rtol=1e-11)     # analytic gradient vs the reference's autodiff
    np.testing.assert_allclose(g["L"] @ g["z"] + float(g["mu"]), g["transform"], rtol=1e-14)
    std = O.laplace_std_from_diag(O.hessian_diag(g["L"], nn, 4, float(g["mu"]), g["z"]))
    np.testing.assert_allclose(std, g["laplace_std"], rtol=1e-6)  # closed form vs the reference's r HVPs


def test_package_stages_match_reference(be):
    g = load("stages")
    X, lm, nn = g["X"], g["landmarks"], g["nn_distances"]
    cov = C.Matern52(float(g["ls"]))
    np.testing.assert_allclose(be.nn_distances(X), nn, rtol=1e-13)
    Lp = mb.parameters.compute_Lp(X, cov, landmarks=lm)
    np.testing.assert_allclose(np.asarray(Lp), g["Lp"], rtol=1e-6, atol=1e-9)
    L = mb.parameters.compute_L(X, cov, landmarks=lm, Lp=Lp)
    np.testing.assert_allclose(np.asarray(L), g["L"], rtol=1e-6, atol=1e-8)
    for tag, rank, gp in (("nystroem15", 15, None), ("nystroem99", 0.99, "sparse_nystroem")):
        Ln = np.asarray(mb.parameters.compute_L(X, cov, landmarks=lm, rank=rank, gp_type=gp))
        assert list(Ln.shape) == list(g[f"L_{tag}_shape"].astype(int))          # rank selection bit-exact
        np.testing.assert_allclose(Ln[:60] @ Ln.T[:, ::5], g[f"L_{tag}_gram"], rtol=1e-6, atol=1e-8)
    Lf = np.asarray(mb.parameters.compute_L(X[:120], cov, rank=0.9, gp_type="full_nystroem"))
    assert list(Lf.shape) == list(g["L_full_nystroem_shape"].astype(int))
    np.testing.assert_allclose(Lf[:60] @ Lf.T[:, ::2], g["L_full_nystroem_gram"], rtol=1e-7, atol=1e-9)
    mu = float(g["mu"])
    Lg = be.upload(g["L"], sharded=True)
    np.testing.assert_allclose(mb.parameters.compute_initial_value(nn, 4, mu, Lg), g["initial_value"], rtol=1e-8)
    transform = mb.inference.compute_transform(mu, Lg)
    loss_func = mb.inference.compute_loss_func(nn, 4, transform, g["z"].shape[0])
    loss, grad = loss_func.value_and_grad(g["z"])
    assert loss == pytest.approx(float(g["loss"]), rel=1e-13)
    np.testing.assert_allclose(grad, g["grad"], rtol=1e-10)
    np.testing.assert_allclose(transform(g["z"]), g["transform"], rtol=1e-13)
    np.testing.assert_allclose(mb.inference.compute_laplace_std(loss_func, g["z"]), g["laplace_std"], rtol=1e-6)


# ---- end to end ----------------------------------------------------------------------------------------------
CASES = {
    # name: (estimator kwargs for the package, oracle kwargs, benign?)
    "config1_full": ({}, {}, True),
    "sparse_uniform_expquad": ({"cov_func_curry": "ExpQuad"}, {"cov_func_curry": "ExpQuad"}, True),
    "sparse_clustered_matern52_laplace": ({"predictor_with_uncertainty": True}, {}, False),
    "nystroem_rank25": ({"rank": 25}, {"rank": 25}, False),
    "nystroem_rank099": ({"rank": 0.99, "gp_type": "sparse_nystroem"},
                         {"rank": 0.99, "gp_type": O.GaussianProcessType.SPARSE_NYSTROEM}, False),
}
# measured reference-vs-reference noise at the default L-BFGS-B stop on the clustered cases (see
# test_estimator_parity.py::test_default_stop_is_within_the_noise_floor); uniform cases meet 1e-5
DEFAULT_STOP_TOL = {True: 1e-5, False: 2e-4}


def _kw(mod, kw):
    kw = dict(kw)
    if "cov_func_curry" in kw:
        kw["cov_func_curry"] = getattr(mod, kw["cov_func_curry"])
    return kw


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_end_to_end_matches_reference(name):
    g = load(name)
    _, okw, benign = CASES[name]
    lm = g["landmarks"] if "landmarks" in g else None
    for tag, opts, tol in (("", None, DEFAULT_STOP_TOL[benign]), ("_tight", TIGHT, 1e-6)):
        fit = O.fit_density(g["X"], landmarks=lm, lbfgsb_options=opts, **_kw(O, okw))
        np.testing.assert_allclose(fit.nn_distances, g["nn_distances"], rtol=1e-13)
        assert fit.mu == pytest.approx(float(g["mu"]), rel=1e-13) and fit.ls == pytest.approx(float(g["ls"]), rel=1e-13)
        assert list(np.shape(fit.L)) == list(g["L_shape"].astype(int))
        assert rel(fit.log_density_x, g["log_density" + tag]) < tol
        assert rel(O.predict_density(fit, g["X"], g["Y"]), g["pred" + tag]) < 10 * tol
        assert abs(fit.loss - float(g["loss" + tag])) < 1e-7 * abs(float(g["loss" + tag]))


@pytest.fixture
def lbfgsb_options():
    old = dict(mb.inference.LBFGSB_OPTIONS)

    def setter(opts):
        mb.inference.LBFGSB_OPTIONS.clear()
        mb.inference.LBFGSB_OPTIONS.update(opts or old)

    yield setter
    setter(old)


@pytest.mark.run_last          # computes its nn_distances itself: on the device since the end of round 1
@pytest.mark.parametrize("name", list(CASES))
def test_package_end_to_end_matches_reference(be, lbfgsb_options, name):
    g = load(name)
    pkw, _, benign = CASES[name]
    lm = g["landmarks"] if "landmarks" in g else None
    for tag, opts, tol in (("", None, DEFAULT_STOP_TOL[benign]), ("_tight", TIGHT, 1e-6)):
        lbfgsb_options(opts)
        est = mb.DensityEstimator(landmarks=lm, **_kw(C, pkw))
        dens = est.fit_predict(g["X"])
        np.testing.assert_allclose(est.nn_distances, g["nn_distances"], rtol=1e-13)
        assert est.mu == pytest.approx(float(g["mu"]), rel=1e-13) and est.ls == pytest.approx(float(g["ls"]), rel=1e-13)
        assert est.gp_type.value == str(g["gp_type"]) and list(est.L.shape) == list(g["L_shape"].astype(int))
        assert type(est.predict).__name__ == str(g["predictor"])
        assert rel(dens, g["log_density" + tag]) < tol
        assert rel(est.predict(g["Y"]), g["pred" + tag]) < 10 * tol
        assert abs(est.losses[-1] - float(g["loss" + tag])) < 1e-7 * abs(float(g["loss" + tag]))
        if "std" in g:
            np.testing.assert_allclose(est.pre_transformation_std, g["std" + tag], rtol=5e-3 if not tag else 1e-4)
            if tag:
                np.testing.assert_allclose(est.predict.covariance(g["Y"]), g["covariance" + tag], rtol=1e-5, atol=1e-9)
                np.testing.assert_allclose(est.predict.mean_covariance(g["Y"]), g["mean_covariance" + tag], rtol=1e-4)


@pytest.mark.run_last          # computes its nn_distances itself: on the device since the end of round 1
def test_package_time_sensitive_matches_reference(be, lbfgsb_options):
    g = load("time_sensitive")
    # converged-optimiser tolerance 3e-6 (not 1e-6): with ftol = 0 L-BFGS-B stops where its line search can
    # no longer resolve the loss, which on this 320-cell problem leaves ~1e-6 of relative play in the log
    # density (measured 1.3e-6 against the JAX reference on B200); north_star's bar is 1e-5.
    for tag, opts, tol in (("", None, 2e-4), ("_tight", TIGHT, 3e-6)):
        lbfgsb_options(opts)
        est = mb.TimeSensitiveDensityEstimator(ls=1.5, ls_time=0.8, landmarks=g["landmarks"])
        dens = est.fit_predict(g["X"], g["times"])
        np.testing.assert_allclose(est.nn_distances, g["nn_distances"], rtol=1e-13)
        assert est.mu == pytest.approx(float(g["mu"]), rel=1e-13)
        assert type(est.predict).__name__ == str(g["predictor"])
        assert rel(dens, g["log_density" + tag]) < tol
        assert rel(est.predict(g["Y"], g["Y_times"]), g["pred" + tag]) < 10 * tol


def test_oracle_time_sensitive_matches_reference():
    g = load("time_sensitive")
    Xt = np.concatenate([g["X"], g["times"][:, None]], axis=1)
    cov = O.Matern52(1.5, active_dims=slice(None, -1)) * O.Matern52(0.8, active_dims=-1)
    fit = O.fit_density(Xt, cov_func=cov, landmarks=g["landmarks"], nn_distances=g["nn_distances"], d=3, ls=1.5,
                        lbfgsb_options=TIGHT)
    assert rel(fit.log_density_x, g["log_density_tight"]) < 1e-6
    Yt = np.concatenate([g["Y"], g["Y_times"][:, None]], axis=1)
    assert rel(O.predict_density(fit, Xt, Yt), g["pred_tight"]) < 1e-5
