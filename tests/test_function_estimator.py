"""FunctionEstimator: GP regression of observed values on the cell states, on the same device
kernels as the density path.

* the oracle restatement (``oracle.mellon_oracle.function_*``) against what the UNMODIFIED reference returns
  (``tests/golden/reference_function_estimator.npz``, minted by ``oracle/make_golden.py --function-only``);
* the package (through the C ABI on the GPU, through the NumPy test double of the ABI without one) against the same
  vectors and against the hard-coded tables of the reference's own golden test (``tests/test_reference_results.py``);
* the API / error contract of the reference's ``tests/test_function_estimator.py``."""

import os

import numpy as np
import pytest

import mellon_b200 as mb
from oracle import mellon_oracle as O

pytestmark = pytest.mark.run_last

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_function_estimator.npz"))
PF = np.array([0.5, 1.0, 2.0])
CASES = {
    "ref_full": dict(sigma=1.0, obs_variance=True),
    "ref_sparse": dict(sigma=1.0, obs_variance=True),
    "full_unc": dict(sigma=0.7, predictor_with_uncertainty=True),
    "sparse_unc": dict(sigma=0.7, predictor_with_uncertainty=True),
    "full_vec": dict(sigma=0.5, obs_variance=True, mu=0.2),
    "sparse_vec": dict(sigma=0.5, obs_variance=True, mu=0.2),
    "full_mean": dict(y_is_mean=True),
    "sparse_mean": dict(y_is_mean=True),
    "full_pf": dict(sigma=PF, obs_variance=True, predictor_with_uncertainty=True),
    "sparse_pf": dict(sigma=PF, obs_variance=True, predictor_with_uncertainty=True),
    "full_np": dict(sigma="np", predictor_with_uncertainty=True),
    "sparse_np": dict(sigma="np", predictor_with_uncertainty=True),
    "full_obs": dict(sigma="obs", predictor_with_uncertainty=True),
    "sparse_obs": dict(sigma="obs", predictor_with_uncertainty=True),
    "clustered_sparse": dict(sigma=0.3, obs_variance=True, cov="Matern32"),
    "clustered_full": dict(sigma=0.3, obs_variance=True),
}
# tests/test_reference_results.py (reference): first rows of its hard-coded tables, atol 1e-5 there and here
REF_TABLES = {
    "ref_full": dict(pred=[[0.1591912, -0.01633006, -0.09774735], [0.22242522, 0.18020723, -0.02099988]],
                     lev=[0.0372332, 0.07869925, 0.12117246, 0.05443739, 0.07560143],
                     obsvar=[[0.95486132, 1.10382589, 1.09700611], [0.99352028, 1.09954301, 1.09154833]]),
    "ref_sparse": dict(pred=[[0.15897022, -0.01638545, -0.09799344], [0.22247079, 0.17997088, -0.02106525]],
                       lev=[0.03717582, 0.07859248, 0.11760941, 0.05433303, 0.07468583],
                       obsvar=[[0.95491038, 1.10365859, 1.0955746], [0.9931193, 1.09942862, 1.09088032]]),
}


SNP = 0.3 + np.random.default_rng(64).random((50, 3))     # the (n, p) sigma of oracle/make_golden.py


def case_kwargs(tag):
    kw = dict(CASES[tag])
    if isinstance(kw.get("sigma"), str):
        kw["sigma"] = SNP if kw["sigma"] == "np" else SNP[:, 0]
    return kw


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def inputs(tag):
    lm = G[tag + "_landmarks"] if tag + "_landmarks" in G else None
    return G[tag + "_X"], G[tag + "_y"], G[tag + "_Xq"], lm


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_reproduces_the_reference(tag):
    kw = case_kwargs(tag)
    cov = getattr(O, kw.pop("cov", "Matern52"))(float(G[tag + "_ls"]))
    mu = kw.pop("mu", 0.0)
    kw["with_uncertainty"] = kw.pop("predictor_with_uncertainty", False)
    X, y, Xq, lm = inputs(tag)
    fit = O.function_fit(X, y, landmarks=lm, mu=mu, cov_func=cov, **kw)
    assert rel(fit.weights, G[tag + "_weights"]) < 1e-9
    assert rel(O.conditional_mean(Xq, fit.base, fit.weights, mu, cov), G[tag + "_pred"]) < 1e-9
    if tag + "_lev" in G:
        assert rel(O.function_leverage(fit, X), G[tag + "_lev"]) < 1e-11
        assert rel(O.function_leverage(fit, Xq), G[tag + "_lev_q"]) < 1e-11
        assert rel(O.function_loo_residuals_squared(fit, X, y), G[tag + "_loo"]) < 1e-11
    if tag + "_obsvar" in G:
        assert rel(fit.corrected_r2, G[tag + "_corrected_r2"]) < 1e-11
        assert rel(fit.variance_weights, G[tag + "_variance_weights"]) < 1e-9
        assert rel(O.function_obs_variance(fit, Xq), G[tag + "_obsvar"]) < 1e-11
    if tag + "_covariance" in G:
        assert rel(O.function_covariance(fit, Xq), G[tag + "_covariance"]) < 1e-10
        assert rel(O.function_covariance(fit, Xq, diag=False), G[tag + "_covariance_full"]) < 1e-10


def fit_package(tag):
    kw = case_kwargs(tag)
    curry = getattr(mb.cov, kw.pop("cov", "Matern52"))
    X, y, Xq, lm = inputs(tag)
    est = mb.FunctionEstimator(cov_func_curry=curry, ls=float(G[tag + "_ls"]), landmarks=lm,
                               n_landmarks=0 if lm is None else None, **kw)
    est.fit(X, y)
    return est, X, y, Xq


def tolerances(be, tag):
    """Relative tolerances (of the largest entry) per quantity.  The NumPy test double shares LAPACK with the stack
    that minted the vectors, so it pins the host logic at rounding level.  On the device the kernel entries differ from
    NumPy's by up to 2e-13 (K1's tolerance), and these quantities amplify that: a 2e-16 perturbation of K moves the
    sparse leverage / variance weights by 1e-10, the sparse weights by 1e-11 (K_uu + 1e-6 I with ls = 4.7 on
    standard-normal cells is conditioned 1e7) and the noise-free full-GP weights by 7e-10 (condition 1e8) — measured
    with the oracle; hence 1e-6 there (1e-5 noise-free), 1e-8 on the well-conditioned predictions."""
    if type(be).__name__ == "FakeBackend":
        return dict(weights=1e-6 if tag == "full_mean" else 1e-9, pred=1e-6 if tag == "full_mean" else 1e-9,
                    lev=1e-9, vw=1e-8, obsvar=1e-9, cov=1e-8)
    noise_free = tag == "full_mean"
    return dict(weights=1e-5 if noise_free else 1e-6, pred=1e-5 if noise_free else 1e-8, lev=1e-6, vw=1e-6,
                obsvar=1e-6, cov=1e-6)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_package_matches_the_reference(be, tag):
    est, X, y, Xq = fit_package(tag)
    pred = est.predict
    tol = tolerances(be, tag)
    assert type(pred).__name__ == str(G[tag + "_predictor"])
    assert rel(pred.weights, G[tag + "_weights"]) < tol["weights"]
    assert rel(pred(Xq), G[tag + "_pred"]) < tol["pred"]
    if tag + "_lev" in G:
        assert rel(pred.leverage(X), G[tag + "_lev"]) < tol["lev"]
        assert rel(est.leverage(), G[tag + "_lev"]) < tol["lev"]
        assert rel(pred.leverage(Xq), G[tag + "_lev_q"]) < tol["lev"]
        assert rel(pred.loo_residuals_squared(X, y), G[tag + "_loo"]) < tol["lev"]
    if tag + "_obsvar" in G:
        assert rel(est.loo_residuals_squared(), G[tag + "_corrected_r2"]) < tol["lev"]
        assert rel(pred.variance_weights, G[tag + "_variance_weights"]) < tol["vw"]
        assert rel(pred.obs_variance(Xq), G[tag + "_obsvar"]) < tol["obsvar"]
        assert rel(est.get_obs_variance(Xq), G[tag + "_obsvar"]) < tol["obsvar"]
    else:
        with pytest.raises(ValueError, match="without obs_variance"):
            pred.obs_variance(Xq)
    if tag + "_covariance" in G:
        nf = dict(noise_free=True) if pred.per_feature_sigma else {}
        assert rel(pred.covariance(Xq, **nf), G[tag + "_covariance"]) < tol["cov"]
        assert rel(pred.covariance(Xq, diag=False, **nf), G[tag + "_covariance_full"]) < tol["cov"]
        if pred.per_feature_sigma:
            with pytest.raises(ValueError, match="noise_free=True"):
                pred.covariance(Xq)
    assert be.launch_count() > 0


@pytest.mark.parametrize("tag", sorted(REF_TABLES))
def test_reference_golden_tables(be, tag):
    """The numbers hard-coded in the reference's tests/test_reference_results.py."""
    table = REF_TABLES[tag]
    est, X, y, Xq = fit_package(tag)
    got = dict(pred=est.predict(Xq)[:2], lev=est.predict.leverage(X)[:5], obsvar=est.predict.obs_variance(Xq)[:2])
    stored = dict(pred=G[tag + "_pred"][:2], lev=G[tag + "_lev"][:5], obsvar=G[tag + "_obsvar"][:2])
    for key, want in table.items():
        assert np.allclose(stored[key], want, atol=1e-5), key     # the stand-in stack reproduces the table
        assert np.allclose(got[key], want, atol=1e-5), key        # and so does the package


@pytest.fixture
def wave():
    rng = np.random.default_rng(535)
    X = rng.multivariate_normal(np.ones(2), [[0.6, 0.2], [0.2, 0.4]], 100)
    clean = np.sum(np.sin(X / 2), axis=1)
    y = clean + 1e-2 * np.sin(1e3 * X).sum(axis=1)
    return X, y, np.stack([y, clean], axis=1), clean


def test_prediction_and_error_contract(be, wave):
    """tests/test_function_estimator.py:24-70 of the reference."""
    X, y, _, clean = wave
    with pytest.raises(ValueError, match="not available for the Function Estimator"):
        mb.FunctionEstimator(gp_type="sparse_nystroem")
    est = mb.FunctionEstimator(sigma=1e-3)
    with pytest.raises(ValueError):
        est.fit_predict()
    pred = est.fit_predict(X, y)
    assert pred.shape == (100,)
    assert len(str(est)) > 0 and "sigma=" in str(est) and len(est._repr_html_()) > 0
    assert np.std(y - pred) < 2e-2 and np.std(clean - pred) < 2e-2
    assert np.allclose(pred, est(X, y))
    est.compute_conditional(y=y)
    est.compute_conditional(x=y, y=y)
    with pytest.raises(ValueError):
        est.compute_conditional(X)
    with pytest.raises(ValueError):
        est.fit(X, y[:3])
    with pytest.raises(ValueError):
        est.fit_predict(X[:, :, None], y)
    with pytest.raises(ValueError):
        est.fit_predict(X[:3, :], y)


def test_multi_output_and_deprecated_form(be, wave):
    X, y, Y, _ = wave
    est = mb.FunctionEstimator(sigma=1e-3)
    both = est.fit_predict(X, Y, X)
    assert both.shape == (100, 2)
    # sigma^2 = jitter = 1e-6: the factor is conditioned ~1e8, so one-column and two-column solves agree to ~1e-8
    assert rel(both[:, 0], mb.FunctionEstimator(sigma=1e-3).fit_predict(X, y)) < 1e-5
    assert rel(mb.FunctionEstimator(sigma=1e-3).multi_fit_predict(X, Y.T, X), both.T) < 1e-9


@pytest.mark.parametrize("n_landmarks, limit", [(0, 1e-4), (10, 1e-1)])
def test_approximations(be, wave, n_landmarks, limit):
    X, y, _, _ = wave
    base = mb.FunctionEstimator(sigma=1e-3).fit_predict(X, y)
    approx = mb.FunctionEstimator(sigma=1e-3, n_landmarks=n_landmarks).fit_predict(X, y)
    assert np.std(approx - base) < limit
    base1 = mb.FunctionEstimator(sigma=1e-3).fit_predict(X[:, 0], y)
    approx1 = mb.FunctionEstimator(sigma=1e-3, n_landmarks=n_landmarks).fit_predict(X[:, 0], y)
    assert np.std(approx1 - base1) < 4e-1


def test_refusals_are_loud(be, wave):
    X, y, Y, _ = wave
    lm = X[:10].copy()
    with pytest.raises(ValueError, match="must be positive"):
        mb.FunctionEstimator(landmarks=lm, ls=1.0).fit(X, y)                       # sigma = 0 divides by zero
    for landmarks in (lm, None):                                                            # (n, p) sigma
        est = mb.FunctionEstimator(landmarks=landmarks, n_landmarks=None if landmarks is not None else 0, ls=1.0,
                                   sigma=np.full_like(Y, 0.1))
        with pytest.raises(ValueError, match="per-observation-per-feature"):
            est.fit(X, Y, obs_variance=True)
        with pytest.raises(ValueError, match="per-observation-per-feature"):
            est.fit(X, Y).predict.leverage(X)
    with pytest.raises(NotImplementedError):
        mb.FunctionEstimator(landmarks=lm, ls=1.0, sigma=np.eye(100) * 0.1).fit(X, y)      # full covariance
    est = mb.FunctionEstimator(landmarks=lm, ls=1.0, sigma=0.1).fit(X, y)
    with pytest.raises(ValueError, match="features"):
        est.predict.leverage(X[:, :1])
    with pytest.raises(ValueError, match="without covariance"):
        est.predict.covariance(X)


def test_predictor_with_observation_variance_round_trips_through_json(be, wave):
    X, y, Y, _ = wave
    est = mb.FunctionEstimator(landmarks=X[:12].copy(), ls=1.2, sigma=0.2, obs_variance=True).fit(X, Y)
    clone = mb.Predictor.from_json_str(est.predict.to_json())
    assert type(clone) is type(est.predict)
    assert rel(clone(X[:7]), est.predict(X[:7])) < 1e-12
    assert rel(clone.obs_variance(X[:7]), est.predict.obs_variance(X[:7])) < 1e-12
    assert rel(clone.leverage(X[:7]), est.predict.leverage(X[:7])) < 1e-12


def test_same_x_rule_follows_the_object_the_caller_passed(be, wave):
    """``set_x`` (base_model.py:176-213): passing the array the estimator was fitted on is fine, another one is an error —
    also when validation had to copy it (an ndarray subclass here; a jax array for the reference's own tests)."""
    X, y, _, _ = wave

    class Tagged(np.ndarray):
        pass

    Xs = X.view(Tagged)
    est = mb.FunctionEstimator(sigma=1e-2, ls=1.0)
    first = est.fit_predict(Xs, y)
    assert type(est.x) is np.ndarray or est.x is Xs
    assert np.allclose(est(Xs, y), first, rtol=1e-10, atol=0)    # same object again: accepted
    with pytest.raises(ValueError, match="has been set already"):
        est.fit_predict(X.copy(), y)                            # equal values, different object: refused as in the reference


def test_random_configurations_against_the_oracle(be):
    """Shapes, noise forms (scalar / per-feature / per-observation / per-observation-per-feature), full and sparse,
    with and without observation variance and covariance: the package against the oracle restatement."""
    rng = np.random.default_rng(0)
    tol = 1e-8 if type(be).__name__ == "FakeBackend" else 1e-6
    for trial in range(24):
        n, d, p = int(rng.integers(20, 90)), int(rng.integers(1, 5)), int(rng.integers(1, 4))
        X, Xq = rng.standard_normal((n, d)), rng.standard_normal((7, d))
        y = rng.standard_normal((n, p)) if rng.random() < 0.7 else rng.standard_normal(n)
        sparse = rng.random() < 0.5
        lm = X[rng.choice(n, int(rng.integers(3, min(15, n - 1))), replace=False)].copy() if sparse else None
        form = rng.choice(["scalar", "pf", "obs", "np"])
        if form == "pf" and y.ndim == 2:
            sigma = rng.random(y.shape[1]) + 0.2
        elif form == "obs" and y.ndim == 1:
            sigma = rng.random(n) + 0.2
        elif form == "np" and y.ndim == 2:
            sigma = rng.random(y.shape) + 0.2
        else:
            sigma, form = float(rng.random() + 0.2), "scalar"
        obs = form in ("scalar", "pf") and rng.random() < 0.6
        unc = rng.random() < 0.5
        ls, mu = float(rng.random() * 2 + 0.5), float(rng.standard_normal())
        est = mb.FunctionEstimator(landmarks=lm, n_landmarks=None if sparse else 0, ls=ls, mu=mu, sigma=sigma,
                                   obs_variance=obs, predictor_with_uncertainty=unc).fit(X, y)
        fit = O.function_fit(X, y, landmarks=lm, mu=mu, cov_func=O.Matern52(ls), sigma=sigma, obs_variance=obs,
                             with_uncertainty=unc)
        case = (trial, n, d, p, sparse, form, obs, unc)
        assert rel(est.predict.weights, fit.weights) < tol, case
        assert rel(est.predict(Xq), O.conditional_mean(Xq, fit.base, fit.weights, mu, fit.cov_func)) < tol, case
        if form in ("scalar", "pf"):
            assert rel(est.predict.leverage(Xq), O.function_leverage(fit, Xq)) < tol, case
        if obs:
            assert rel(est.predict.obs_variance(Xq), O.function_obs_variance(fit, Xq)) < tol, case
        if unc:
            nf = dict(noise_free=True) if est.predict.per_feature_sigma else {}
            assert rel(est.predict.covariance(Xq, **nf), O.function_covariance(fit, Xq)) < tol, case


FUNCTION_PREDICTORS = os.path.join(os.path.dirname(__file__), "golden", "reference_function_predictors.json")


@pytest.mark.parametrize("name", ["function_full", "function_sparse"])
def test_reference_written_function_predictor_loads_and_predicts(be, name):
    """Predictors fitted and serialised by the unmodified reference (oracle/make_golden.py --function-predictors-only):
    the JSON text loads here and gives the reference's own numbers; the reverse direction is tools/check_json_interop.py."""
    import json

    with open(FUNCTION_PREDICTORS) as f:
        case = json.load(f)[name]
    pred = mb.Predictor.from_json_str(case["json"])
    assert type(pred).__name__ == case["classname"] and bool(pred.per_feature_sigma) == case["per_feature_sigma"]
    Y = np.asarray(case["Y"])
    nf = dict(noise_free=True) if pred.per_feature_sigma else {}
    tol = 1e-9 if type(be).__name__ == "FakeBackend" else 1e-6
    assert rel(pred(Y), case["mean"]) < tol
    assert rel(pred.leverage(Y), case["leverage"]) < tol
    assert rel(pred.obs_variance(Y), case["obs_variance"]) < tol
    assert rel(pred.covariance(Y, **nf), case["covariance"]) < tol
