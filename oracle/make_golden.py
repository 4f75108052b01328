#!/usr/bin/env python
"""Mint the golden vectors of tests/golden/ by running the UNMODIFIED reference
(/root/reference/mellon, imported from where it lies) on the NumPy stand-ins of oracle/refshim.

    python oracle/make_golden.py            # writes tests/golden/reference_*.npz

Inputs are seeded NumPy arrays (stored in the files); outputs are what the reference's own code
returns.  Each end-to-end case is stored twice: at the reference's default L-BFGS-B stop and with
the optimiser run to convergence (``*_tight``), because the default stop carries a ~1e-5
reference-vs-reference noise floor on clustered data ("Parity").  This script is the
only thing that reads /root/reference; the tests read the committed .npz files."""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, os.environ.get("MELLON_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

import importlib  # noqa: E402

import jaxopt  # noqa: E402  (the stand-in)
import mellon  # noqa: E402  (the reference)

P = importlib.import_module("mellon.parameters")   # mellon/parameters.py itself (mellon.parameters is the re-export shim)
INF = importlib.import_module("mellon.inference")

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
TIGHT = dict(maxiter=20000, maxfun=100000, ftol=0.0, gtol=1e-9)
mellon.setup_logging().setLevel("WARNING")


def blobs(n, d, seed, k=6, spread=0.25):
    centers = np.random.default_rng(1000 + d).standard_normal((k, d)) * 1.5
    rng = np.random.default_rng(seed)
    return centers[rng.integers(0, k, n)] + spread * rng.standard_normal((n, d))


def A(x):
    return np.asarray(x, dtype=float)


def fit_case(make_est, X, Y, fit_args=(), pred_args=(), with_unc=False):
    out = {}
    for tag, opts in (("", {}), ("_tight", TIGHT)):
        jaxopt.EXTRA_OPTIONS.clear()
        jaxopt.EXTRA_OPTIONS.update(opts)
        est = make_est()
        dens = est.fit_predict(X, *fit_args)
        out["log_density" + tag] = A(dens)
        out["pre_transformation" + tag] = A(est.pre_transformation)
        out["loss" + tag] = A(est.losses[-1])
        out["nfev" + tag] = A(est.opt_state.num_fun_eval)
        out["pred" + tag] = A(est.predict(Y, *pred_args))
        if with_unc:
            out["std" + tag] = A(est.pre_transformation_std)
            out["covariance" + tag] = A(est.predict.covariance(Y, *pred_args))
            out["mean_covariance" + tag] = A(est.predict.mean_covariance(Y, *pred_args))
    jaxopt.EXTRA_OPTIONS.clear()
    out.update(mu=A(est.mu), ls=A(est.ls), d=A(est.d), nn_distances=A(est.nn_distances),
               initial_value=A(est.initial_value), L_shape=A(est.L.shape), gp_type=np.array(str(est.gp_type.value)),
               predictor=np.array(type(est.predict).__name__))
    if est.landmarks is not None:
        out["landmarks"] = A(est.landmarks)
    return est, out


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"reference_{name}.npz")
    np.savez_compressed(path, **arrays)
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB, keys={sorted(arrays)}")


def main():
    rng = np.random.default_rng(7)

    # ---- kernels (cov.py, base_cov.py, util.distance) -------------------------------------------
    x, y = rng.random((23, 5)), rng.random((19, 5))
    kern = {"x": x, "y": y}
    C = mellon.cov
    for name in ("Matern32", "Matern52", "ExpQuad", "Exponential", "Linear"):
        kern[name] = A(getattr(C, name)(1.3)(x, y))
        kern[name + "_dims"] = A(getattr(C, name)(0.7, active_dims=[0, 3])(x, y))
    kern["RatQuad"] = A(C.RatQuad(2.5, 1.3)(x, y))
    kern["distance"] = A(mellon.util.distance(x, y))
    kern["diag_Matern52"] = A(C.Matern52(1.3).diag(x))
    expr = (C.Matern52(1.2, active_dims=slice(None, -1)) * C.ExpQuad(0.4, active_dims=-1) + 0.3) ** 2
    kern["expr"] = A(expr(x, y))
    nested = C.Matern32(0.7, active_dims=[0, 2]) + C.ExpQuad(1.1, active_dims=1)
    nested.active_dims = [1, 3, 4]
    kern["nested"] = A(nested(x, y))
    save("kernels", **kern)

    # ---- decompositions and the objective at fixed inputs -----------------------------------------
    X = blobs(400, 4, 11)
    lm = X[:40].copy()
    nn = A(P.compute_nn_distances(X))
    cov = C.Matern52(2.0)
    Lp = A(P.compute_Lp(X, cov, landmarks=lm))
    L = A(P.compute_L(X, cov, landmarks=lm, Lp=Lp))
    Lny = A(P.compute_L(X, cov, landmarks=lm, rank=15))
    Lny99 = A(P.compute_L(X, cov, landmarks=lm, rank=0.99, gp_type="sparse_nystroem"))
    Lfn = A(P.compute_L(X[:120], cov, rank=0.9, gp_type="full_nystroem"))
    d = 4
    mu = P.compute_mu(nn, d)
    z0 = A(P.compute_initial_value(nn, d, mu, L))
    transform = INF.compute_transform(mu, L)
    loss_func = INF.compute_loss_func(nn, d, transform, z0.shape[0])
    z = np.random.default_rng(3).standard_normal(z0.shape[0]) * 0.3
    import jax

    val, grad = jax.value_and_grad(loss_func)(z)
    std = A(INF.compute_laplace_std(loss_func, z))
    save("stages", X=X, landmarks=lm, nn_distances=nn, ls=A(2.0), Lp=Lp, L=L, L_nystroem15_gram=Lny[:60] @ Lny.T[:, ::5],
         L_nystroem15_shape=A(Lny.shape), L_nystroem99_shape=A(Lny99.shape), L_nystroem99_gram=Lny99[:60] @ Lny99.T[:, ::5],
         L_full_nystroem_shape=A(Lfn.shape), L_full_nystroem_gram=Lfn[:60] @ Lfn.T[:, ::2], mu=A(mu),
         ls_heuristic=A(P.compute_ls(nn)), initial_value=z0, z=z, loss=A(val), grad=A(grad),
         laplace_std=std, transform=A(transform(z)))

    # ---- end to end ----------------------------------------------------------------------------------
    # config 1 of BASELINE.json: README smoke test, FULL GP
    X1 = np.random.default_rng(0).random((100, 10))
    Y1 = np.random.default_rng(1).random((100, 10))
    _, out = fit_case(lambda: mellon.DensityEstimator(), X1, Y1)
    save("config1_full", X=X1, Y=Y1, **out)

    # sparse Cholesky on uniform cells (the BASELINE synthetic shape, scaled down) and on clusters
    Xu = np.random.default_rng(0).random((600, 20))
    Yu = np.random.default_rng(2).random((50, 20))
    lmu = Xu[:60].copy()
    _, out = fit_case(lambda: mellon.DensityEstimator(cov_func_curry=C.ExpQuad, landmarks=lmu), Xu, Yu)
    save("sparse_uniform_expquad", X=Xu, Y=Yu, **out)
    Xc, Yc = blobs(600, 5, 21), blobs(50, 5, 22)
    lmc = Xc[:60].copy()
    _, out = fit_case(lambda: mellon.DensityEstimator(landmarks=lmc, predictor_with_uncertainty=True), Xc, Yc,
                      with_unc=True)
    save("sparse_clustered_matern52_laplace", X=Xc, Y=Yc, **out)
    _, out = fit_case(lambda: mellon.DensityEstimator(landmarks=lmc, rank=25), Xc, Yc)
    save("nystroem_rank25", X=Xc, Y=Yc, **out)
    _, out = fit_case(lambda: mellon.DensityEstimator(landmarks=lmc, rank=0.99, gp_type="sparse_nystroem"), Xc, Yc)
    save("nystroem_rank099", X=Xc, Y=Yc, **out)

    # time-sensitive: 4 time points x 80 cells, Matern52 x Matern52(time)
    Xt = np.concatenate([blobs(80, 3, 30 + t) + 0.2 * t for t in range(4)])
    times = np.repeat(np.arange(4.0), 80)
    lmt = np.concatenate([Xt, times[:, None]], axis=1)[::8].copy()
    _, out = fit_case(lambda: mellon.TimeSensitiveDensityEstimator(ls=1.5, ls_time=0.8, landmarks=lmt), Xt, Xt[:30],
                      fit_args=(times,), pred_args=(times[:30],))
    save("time_sensitive", X=Xt, times=times, Y=Xt[:30], Y_times=times[:30], **out)

    serialised_predictors()
    function_estimator()
    serialised_function_predictors()


def serialised_predictors():
    """Predictors fitted AND serialised by the reference (``Predictor.to_json``, base_predictor.py:541-734), with the
    reference's own predictions at a few query points: tests/test_serialisation_interop.py loads the JSON text with
    this package and must reproduce them."""
    import json

    C = mellon.cov
    cases = {}
    Xc, Yc = blobs(300, 4, 41), blobs(12, 4, 42)
    lmc = Xc[:30].copy()
    X1 = np.random.default_rng(5).random((60, 3))
    Xt = np.concatenate([blobs(50, 2, 50 + t) + 0.2 * t for t in range(3)])
    times = np.repeat(np.arange(3.0), 50)
    lmt = np.concatenate([Xt, times[:, None]], axis=1)[::6].copy()
    todo = {
        "sparse_cholesky_laplace": (lambda: mellon.DensityEstimator(landmarks=lmc, predictor_with_uncertainty=True), Xc, (), Yc, ()),
        "sparse_nystroem": (lambda: mellon.DensityEstimator(landmarks=lmc, rank=12), Xc, (), Yc, ()),
        "full": (lambda: mellon.DensityEstimator(cov_func_curry=C.ExpQuad), X1, (), X1[:9] + 0.01, ()),
        "time_sensitive": (lambda: mellon.TimeSensitiveDensityEstimator(ls=1.5, ls_time=0.8, landmarks=lmt), Xt, (times,), Xt[:10], (times[:10],)),
    }
    for name, (make, X, fit_args, Y, pred_args) in todo.items():
        est = make()
        est.fit(X, *fit_args)
        pred = est.predict
        case = {"json": pred.to_json(), "Y": A(Y).tolist(), "pred_args": [A(a).tolist() for a in pred_args],
                "mean": A(pred(Y, *pred_args)).tolist(), "mean_normalized": A(pred(Y, *pred_args, normalize=True)).tolist(),
                "classname": type(pred).__name__, "cov_func_json": est.cov_func.to_json()}
        if name == "sparse_cholesky_laplace":
            case["covariance"] = A(pred.covariance(Y)).tolist()
            case["mean_covariance"] = A(pred.mean_covariance(Y)).tolist()
            case["uncertainty"] = A(pred.uncertainty(Y)).tolist()
        if name == "time_sensitive":
            case["mean_multi_time"] = A(pred(Y[:, :], multi_time=[0.0, 1.5])).tolist()
        cases[name] = case
    path = os.path.join(OUT, "reference_predictors.json")
    with open(path, "w") as f:
        json.dump(cases, f)
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB, cases={sorted(cases)}")


def function_estimator():
    """FunctionEstimator (function_estimator.py:29-615) run by the reference: the two cases of the
    reference's own golden test (tests/test_reference_results.py — inputs from ``jax.random.PRNGKey(42)`` through the
    threefry restatement; the hard-coded tables of that file are asserted in tests/test_function_estimator.py), plus
    a vector ``y``, ``y_is_mean``, per-feature sigma, predictive covariance and a clustered 400-cell problem."""
    import jax

    out = {}

    def run(tag, X, y, Xq, landmarks=None, n_landmarks=None, no_leverage=False, **kw):
        est = mellon.FunctionEstimator(landmarks=landmarks, n_landmarks=n_landmarks, **kw)
        est.fit(X, y)
        pred = est.predict
        out[tag + "_X"], out[tag + "_y"], out[tag + "_Xq"] = A(X), A(y), A(Xq)
        out[tag + "_ls"] = A(est.ls)
        out[tag + "_predictor"] = np.array(type(pred).__name__)
        if est.landmarks is not None:
            out[tag + "_landmarks"] = A(est.landmarks)
        out[tag + "_pred"] = A(pred(Xq))
        out[tag + "_weights"] = A(pred.weights)
        if not kw.get("y_is_mean", False) and not no_leverage:
            out[tag + "_lev"] = A(pred.leverage(X))
            out[tag + "_lev_q"] = A(pred.leverage(Xq))
            out[tag + "_loo"] = A(pred.loo_residuals_squared(X, y))
        if kw.get("obs_variance", False):
            out[tag + "_obsvar"] = A(pred.obs_variance(Xq))
            out[tag + "_variance_weights"] = A(pred.variance_weights)
            out[tag + "_corrected_r2"] = A(est.loo_residuals_squared())
        if kw.get("predictor_with_uncertainty", False):
            nf = dict(noise_free=True) if pred.per_feature_sigma else {}
            out[tag + "_covariance"] = A(pred.covariance(Xq, **nf))
            out[tag + "_covariance_full"] = A(pred.covariance(Xq, diag=False, **nf))
        return est

    k1, k2, k3 = jax.random.split(jax.random.PRNGKey(42), 3)
    X, y, Xq = A(jax.random.normal(k1, (50, 2))), A(jax.random.normal(k2, (50, 3))), A(jax.random.normal(k3, (10, 2)))
    run("ref_full", X, y, Xq, n_landmarks=0, sigma=1.0, obs_variance=True)
    sp = run("ref_sparse", X, y, Xq, n_landmarks=15, sigma=1.0, obs_variance=True)
    lm = A(sp.landmarks)
    run("full_unc", X, y, Xq, n_landmarks=0, sigma=0.7, predictor_with_uncertainty=True)
    run("sparse_unc", X, y, Xq, landmarks=lm, sigma=0.7, predictor_with_uncertainty=True)
    run("full_vec", X, y[:, 0], Xq, n_landmarks=0, sigma=0.5, obs_variance=True, mu=0.2)
    run("sparse_vec", X, y[:, 0], Xq, landmarks=lm, sigma=0.5, obs_variance=True, mu=0.2)
    run("full_mean", X, y, Xq, n_landmarks=0, y_is_mean=True)
    run("sparse_mean", X, y[:, 1], Xq, landmarks=lm, y_is_mean=True)
    pf = np.array([0.5, 1.0, 2.0])
    run("full_pf", X, y, Xq, n_landmarks=0, sigma=pf, obs_variance=True, predictor_with_uncertainty=True)
    run("sparse_pf", X, y, Xq, landmarks=lm, sigma=pf, obs_variance=True, predictor_with_uncertainty=True)
    # one noise level per observation and output, and per observation for a vector y (tests/test_perobservation_sigma.py)
    snp = 0.3 + np.random.default_rng(64).random((50, 3))
    run("full_np", X, y, Xq, n_landmarks=0, sigma=snp, predictor_with_uncertainty=True, no_leverage=True)
    run("sparse_np", X, y, Xq, landmarks=lm, sigma=snp, predictor_with_uncertainty=True, no_leverage=True)
    run("full_obs", X, y[:, 2], Xq, n_landmarks=0, sigma=snp[:, 0], predictor_with_uncertainty=True, no_leverage=True)
    run("sparse_obs", X, y[:, 2], Xq, landmarks=lm, sigma=snp[:, 0], predictor_with_uncertainty=True, no_leverage=True)
    Xc, Xcq = blobs(400, 4, 61), blobs(30, 4, 62)
    yc = np.stack([np.sin(Xc[:, 0]) + 0.1 * np.random.default_rng(63).standard_normal(400), Xc[:, 1] * Xc[:, 2]], axis=1)
    run("clustered_sparse", Xc, yc, Xcq, landmarks=Xc[:40].copy(), sigma=0.3, obs_variance=True, ls=1.5,
        cov_func_curry=mellon.cov.Matern32)
    run("clustered_full", Xc[:150], yc[:150], Xcq, n_landmarks=0, sigma=0.3, obs_variance=True, ls=1.5)
    save("function_estimator", **out)


def serialised_function_predictors():
    """FunctionEstimator predictors fitted AND serialised by the reference, with its own numbers at a few query points:
    tests/test_function_estimator.py loads the JSON text with this package (SURVEY.md §8f.3 + §8f.4)."""
    import json

    Xc, Yc = blobs(300, 4, 41), blobs(12, 4, 42)
    yc = np.stack([np.sin(Xc[:, 0]), Xc[:, 1] * Xc[:, 2]], axis=1)
    todo = {
        "function_sparse": (dict(landmarks=Xc[:30].copy(), ls=1.5, sigma=np.array([0.3, 0.6]), obs_variance=True,
                                 predictor_with_uncertainty=True), Xc, yc),
        "function_full": (dict(n_landmarks=0, ls=1.5, sigma=0.4, obs_variance=True, predictor_with_uncertainty=True),
                          Xc[:60], yc[:60]),
    }
    cases = {}
    for name, (kw, X, y) in todo.items():
        est = mellon.FunctionEstimator(**kw)
        est.fit(X, y)
        pred = est.predict
        nf = dict(noise_free=True) if pred.per_feature_sigma else {}
        cases[name] = {"json": pred.to_json(), "Y": A(Yc).tolist(), "classname": type(pred).__name__,
                       "mean": A(pred(Yc)).tolist(), "leverage": A(pred.leverage(Yc)).tolist(),
                       "obs_variance": A(pred.obs_variance(Yc)).tolist(), "covariance": A(pred.covariance(Yc, **nf)).tolist(),
                       "per_feature_sigma": bool(pred.per_feature_sigma)}
    path = os.path.join(OUT, "reference_function_predictors.json")
    with open(path, "w") as f:
        json.dump(cases, f)
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB, cases={sorted(cases)}")


if __name__ == "__main__":
    if "--function-predictors-only" in sys.argv:
        serialised_function_predictors()
    elif "--predictors-only" in sys.argv:
        serialised_predictors()
    elif "--function-only" in sys.argv:
        function_estimator()
    else:
        main()
