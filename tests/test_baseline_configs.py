"""Parity at BASELINE.json's configuration SHAPES (M = 5000 landmarks, D = 50 / 20, the configured kernels) at cell
counts the CPU oracle finishes in about a minute.  Each case asserts three things, with the reference's own
acceptance metric std(a - b) / std(b) (tests/test_density_estimator.py:30-44) for the log densities:

  1. stage-wise, at fixed inputs: the factor Lp reproduces K_MM + jitter I to rounding (backward error), L and the
     Ridge start z0 agree with the oracle's to the accuracy the conditioning of the problem allows;
  2. converged: with both optimisers run to convergence (gtol 1e-9) the log densities agree to 1e-6 — same objective,
     same optimum — or, where the optimum is so flat that the oracle does not reproduce ITSELF to 1e-6 at that
     tolerance (landmarks permuted), to three times that measured floor;
  3. default stop (SciPy's ftol, what the reference runs): the gap to the oracle is no larger than the gap between the
     oracle and ITSELF when its landmarks are permuted — identical mathematics, different rounding.  That
     reference-vs-reference floor is measured in the test; where it is below 1e-5 the 1e-5 bar of north_star applies.

GPU only: on the NumPy test double these would compare NumPy with NumPy."""

import numpy as np
import pytest

import mellon_b200 as mb
from mellon_b200 import cov as C
from oracle import mellon_oracle as O

TIGHT = dict(maxiter=20000, maxfun=100000, ftol=0.0, gtol=1e-9)


def relstd(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.std(a - b) / np.std(b))


def landmarks_of(x, m, seed=1):
    return np.ascontiguousarray(x[np.sort(np.random.default_rng(seed).choice(x.shape[0], m, replace=False))])


@pytest.fixture
def cuda(be):
    if be.name != "cuda":
        pytest.skip("parity of the CUDA library at BASELINE shapes (the test double is NumPy itself)")
    return be


def with_options(opts, fn):
    old = dict(mb.inference.LBFGSB_OPTIONS)
    mb.inference.LBFGSB_OPTIONS.clear()
    mb.inference.LBFGSB_OPTIONS.update(opts)
    try:
        return fn()
    finally:
        mb.inference.LBFGSB_OPTIONS.clear()
        mb.inference.LBFGSB_OPTIONS.update(old)


def check_density_config(cuda, X, lm, nn, cov_o, cov_c, rank=None, label=""):
    kw = dict(landmarks=lm, nn_distances=nn, rank=rank)
    ref = O.fit_density(X, cov_func_curry=cov_o, **kw)
    perm = np.random.default_rng(7).permutation(lm.shape[0])
    ref_perm = O.fit_density(X, cov_func_curry=cov_o, landmarks=np.ascontiguousarray(lm[perm]), nn_distances=nn, rank=rank)
    best = O.fit_density(X, cov_func_curry=cov_o, lbfgsb_options=TIGHT, Lp=ref.Lp if rank is None else None,
                         L=ref.L if rank is None else None, **kw)
    best_perm = O.fit_density(X, cov_func_curry=cov_o, landmarks=np.ascontiguousarray(lm[perm]), nn_distances=nn, rank=rank,
                              lbfgsb_options=TIGHT, Lp=ref_perm.Lp if rank is None else None,
                              L=ref_perm.L if rank is None else None)
    floor = relstd(ref_perm.log_density_x, ref.log_density_x)
    floor_conv = relstd(best_perm.log_density_x, best.log_density_x)

    est = mb.DensityEstimator(cov_func_curry=cov_c, check_rank=False, **kw)
    dens = est.fit_predict(X)
    # 1. stage-wise
    if rank is None:
        Lp, L = np.asarray(est.Lp), np.asarray(est.L)
        W = cov_o(ref.ls)(lm, lm) + 1e-6 * np.eye(lm.shape[0])
        back = np.max(np.abs(Lp @ Lp.T - W)) / np.max(np.abs(W))
        back_ref = np.max(np.abs(ref.Lp @ ref.Lp.T - W)) / np.max(np.abs(W))
        assert back < 1e-14 + 4 * back_ref, (back, back_ref)                 # as good a factor as LAPACK's
        dL = float(np.max(np.abs(L - ref.L)))
        # forward differences of two backward-stable factorisations of a matrix with condition ~1e8 .. 1e10
        assert dL < 1e-7, dL
        resid = np.max(np.abs(L[:2000] @ Lp.T - cov_o(ref.ls)(X[:2000], lm)))   # L Lp^T = K_NM to rounding
        assert resid < 1e-13, resid
        z0 = np.asarray(est.initial_value)
        dz0 = float(np.max(np.abs(z0 - ref.initial_value)) / np.max(np.abs(ref.initial_value)))
        assert dz0 < 1e-5, dz0
    else:
        assert est.L.shape == ref.L.shape                                        # integer rank selection is exact
        dL = dz0 = float("nan")
    # 2. converged
    dens_tight = with_options(TIGHT, lambda: mb.DensityEstimator(cov_func_curry=cov_c, check_rank=False, **kw).fit_predict(X))
    conv = relstd(dens_tight, best.log_density_x)
    # 3. default stop against the measured reference-vs-reference floor
    gap = relstd(dens, ref.log_density_x)
    print(f"\n[{label}] N={X.shape[0]} M={lm.shape[0]}: |dL|={dL:.2e} |dz0|/|z0|={dz0:.2e} converged {conv:.2e} "
          f"(oracle-vs-oracle {floor_conv:.2e}) default-stop gap {gap:.2e} (oracle-vs-oracle {floor:.2e}) nfev cuda/oracle "
          f"{est.opt_state.num_fun_eval}/{ref.opt_state.num_fun_eval}")
    assert conv < max(1e-6, 3 * floor_conv), (conv, floor_conv)
    assert gap < max(1e-5, 3 * floor), (gap, floor)
    return est, ref


def test_config2_shape_expquad_m5000(cuda):
    """BASELINE configs[1] shape: D = 50, M = 5000, ExpQuad, sparse Cholesky (N = 30 000 of the 100 000)."""
    X = np.random.default_rng(0).random((30_000, 50))
    lm = landmarks_of(X, 5000)
    nn = cuda.nn_distances(X)
    est, ref = check_density_config(cuda, X, lm, nn, O.ExpQuad, C.ExpQuad, label="config 2")
    # configs[4] shape: out-of-sample predict on the fitted model.  The weights Lp^-T z amplify the default-stop
    # difference of z by 1 / sqrt(jitter), so the predictions are compared at the oracle's own pre_transformation
    Y = np.random.default_rng(2).random((20_000, 50))
    pred_fn = mb.conditional.LandmarksConditionalCholesky(lm, ref.pre_transformation, ref.mu, est.cov_func, X.shape[0],
                                                          L=est.Lp)
    gap_pred = relstd(pred_fn(Y), O.predict_density(ref, X, Y))
    print(f"[config 5] predict 20 000 queries with the oracle's latent vector: {gap_pred:.2e}")
    assert gap_pred < 1e-5


def test_config3_shape_matern52_nystroem_rank2000(cuda):
    """BASELINE configs[2] shape: D = 50, M = 5000, Matern52, Nystroem rank = 2000 (N = 20 000)."""
    X = np.random.default_rng(0).random((20_000, 50))
    lm = landmarks_of(X, 5000)
    nn = cuda.nn_distances(X)
    check_density_config(cuda, X, lm, nn, O.Matern52, C.Matern52, rank=2000, label="config 3")


def test_config4_shape_time_sensitive(cuda):
    """BASELINE configs[3] shape: D = 20, 10 time points, Matern32 (state) x ExpQuad (time), N = 20 000, M = 2000."""
    n, ls, ls_time = 20_000, 6.0, 1.5
    X = np.random.default_rng(0).random((n, 20))
    times = np.repeat(np.arange(10.0), n // 10)
    Xt = np.concatenate([X, times[:, None]], axis=1)
    lm = landmarks_of(Xt, 2000)
    cov_c = C.Matern32(ls, active_dims=slice(None, -1)) * C.ExpQuad(ls_time, active_dims=-1)
    cov_o = O.Matern32(ls, active_dims=slice(None, -1)) * O.ExpQuad(ls_time, active_dims=-1)
    est = mb.TimeSensitiveDensityEstimator(cov_func=cov_c, ls=ls, ls_time=ls_time, landmarks=lm, check_rank=False)
    dens = est.fit_predict(X, times)
    nn = np.asarray(est.nn_distances)
    kw = dict(cov_func=cov_o, landmarks=lm, nn_distances=nn, d=20, ls=ls)
    ref = O.fit_density(Xt, **kw)
    perm = np.random.default_rng(7).permutation(lm.shape[0])
    ref_perm = O.fit_density(Xt, cov_func=cov_o, landmarks=np.ascontiguousarray(lm[perm]), nn_distances=nn, d=20, ls=ls)
    best = O.fit_density(Xt, lbfgsb_options=TIGHT, Lp=ref.Lp, L=ref.L, **kw)
    floor = relstd(ref_perm.log_density_x, ref.log_density_x)
    dL = float(np.max(np.abs(np.asarray(est.L) - ref.L)))
    dz0 = float(np.max(np.abs(np.asarray(est.initial_value) - ref.initial_value)) / np.max(np.abs(ref.initial_value)))
    dens_tight = with_options(TIGHT, lambda: mb.TimeSensitiveDensityEstimator(
        cov_func=cov_c, ls=ls, ls_time=ls_time, landmarks=lm, nn_distances=nn, check_rank=False).fit_predict(X, times))
    conv, gap = relstd(dens_tight, best.log_density_x), relstd(dens, ref.log_density_x)
    print(f"\n[config 4] N={n} M=2000: |dL|={dL:.2e} |dz0|/|z0|={dz0:.2e} converged {conv:.2e} default-stop gap {gap:.2e} "
          f"(oracle-vs-oracle floor {floor:.2e}) nfev cuda/oracle {est.opt_state.num_fun_eval}/{ref.opt_state.num_fun_eval}")
    assert dL < 1e-7 and dz0 < 1e-5
    assert conv < 1e-6, conv
    assert gap < max(1e-5, 3 * floor), (gap, floor)
