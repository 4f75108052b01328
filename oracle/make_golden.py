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


if __name__ == "__main__":
    if "--predictors-only" in sys.argv:
        serialised_predictors()
    else:
        main()
