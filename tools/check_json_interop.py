"""Reverse direction of tests/test_serialisation_interop.py: predictors fitted and written by THIS package are loaded
by the unmodified reference (imported from /root/reference on the NumPy stand-ins of oracle/refshim) and must predict
the same numbers there.  Local tool (reads /root/reference); output kept as profiles/json_interop_r01.txt.

    python tools/check_json_interop.py [--cuda]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(0, os.environ.get("MELLON_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

import mellon  # noqa: E402  (the reference)
import mellon_b200 as mb  # noqa: E402

if "--cuda" not in sys.argv:
    from fake_lib import FakeBackend  # noqa: E402

    mb.set_backend(FakeBackend(0, 1))
mellon.setup_logging().setLevel("ERROR")
mb.setup_logging().setLevel("ERROR")


def blobs(n, d, seed, k=6, spread=0.25):
    centers = np.random.default_rng(1000 + d).standard_normal((k, d)) * 1.5
    rng = np.random.default_rng(seed)
    return centers[rng.integers(0, k, n)] + spread * rng.standard_normal((n, d))


Xc, Yc = blobs(300, 4, 41), blobs(12, 4, 42)
lmc = Xc[:30].copy()
X1 = np.random.default_rng(5).random((60, 3))
Xt = np.concatenate([blobs(50, 2, 50 + t) + 0.2 * t for t in range(3)])
times = np.repeat(np.arange(3.0), 50)
lmt = np.concatenate([Xt, times[:, None]], axis=1)[::6].copy()
cases = {
    "sparse_cholesky_laplace": (lambda: mb.DensityEstimator(landmarks=lmc, predictor_with_uncertainty=True), Xc, (), Yc, ()),
    "sparse_nystroem": (lambda: mb.DensityEstimator(landmarks=lmc, rank=12), Xc, (), Yc, ()),
    "full": (lambda: mb.DensityEstimator(cov_func_curry=mb.cov.ExpQuad), X1, (), X1[:9] + 0.01, ()),
    "time_sensitive": (lambda: mb.TimeSensitiveDensityEstimator(ls=1.5, ls_time=0.8, landmarks=lmt), Xt, (times,), Xt[:10], (times[:10],)),
    "composite_cov": (lambda: mb.DensityEstimator(cov_func=(mb.cov.Matern32(1.1, active_dims=[0, 1]) + mb.cov.RatQuad(2.0, 0.9)) * 0.7 + 0.05,
                                                  landmarks=lmc), Xc, (), Yc, ()),
}
# regression predictors (FunctionEstimator): multi-output weights, per-feature sigma, observation-variance weights
yc = np.stack([np.sin(Xc[:, 0]), Xc[:, 1] * Xc[:, 2]], axis=1)
cases["function_sparse_obsvar"] = (lambda: mb.FunctionEstimator(landmarks=lmc, ls=1.5, sigma=np.array([0.3, 0.6]), obs_variance=True,
                                                               predictor_with_uncertainty=True), Xc, (yc,), Yc, ())
cases["function_full_obsvar"] = (lambda: mb.FunctionEstimator(n_landmarks=0, ls=1.5, sigma=0.4, obs_variance=True,
                                                             predictor_with_uncertainty=True), Xc[:120], (yc[:120],), Yc, ())
worst = 0.0
for name, (make, X, fit_args, Y, pred_args) in cases.items():
    est = make()
    est.fit(X, *fit_args)
    ours = est.predict
    theirs = mellon.Predictor.from_json_str(ours.to_json())
    assert type(theirs).__module__.startswith("mellon.") and type(theirs).__name__ == type(ours).__name__
    checks = {"mean": (ours(Y, *pred_args), theirs(Y, *pred_args))}
    if name.startswith("function_"):
        nf = dict(noise_free=True) if ours.per_feature_sigma else {}
        checks["obs_variance"] = (ours.obs_variance(Y), theirs.obs_variance(Y))
        checks["leverage"] = (ours.leverage(Y), theirs.leverage(Y))
        checks["covariance"] = (ours.covariance(Y, **nf), theirs.covariance(Y, **nf))
    else:
        checks["mean normalized"] = (ours(Y, *pred_args, normalize=True), theirs(Y, *pred_args, normalize=True))
    if name == "sparse_cholesky_laplace":
        checks["covariance"] = (ours.covariance(Y), theirs.covariance(Y))
        checks["mean_covariance"] = (ours.mean_covariance(Y), theirs.mean_covariance(Y))
    for what, (a, b) in checks.items():
        a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
        err = float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))
        worst = max(worst, err)
        print(f"{name:26s} {type(theirs).__name__:30s} {what:16s} max|ours - reference-loaded| / max|.| = {err:.2e}")
    # and a covariance document alone
    back = mellon.cov.Covariance.from_json(est.cov_func.to_json())
    lmk = np.asarray(ours.landmarks if getattr(ours, "landmarks", None) is not None else ours.x)
    k_r = np.asarray(back(lmk[:5], lmk[:7]))
    k_o = np.asarray(est.cov_func(lmk[:5], lmk[:7]))
    err = float(np.max(np.abs(k_o - k_r)))
    worst = max(worst, err)
    print(f"{name:26s} {type(back).__name__:30s} {'cov_func':16s} max|K ours - K reference-loaded| = {err:.2e}")
print(f"worst = {worst:.2e}  ->", "OK" if worst < 1e-9 else "MISMATCH")
sys.exit(0 if worst < 1e-9 else 1)
