"""Config-4-shaped (time-sensitive, Matern32 x ExpQuad(time)) parity: default L-BFGS-B stop vs converged, and the
oracle's own rounding-level sensitivity (landmarks permuted)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mellon_b200 as mb
from mellon_b200 import cov as C
from oracle import mellon_oracle as O

mb.setup_logging().setLevel("WARNING")
be = mb.get_backend()
n4 = int(os.environ.get("PARITY_N4", 20000))
X4 = np.random.default_rng(0).random((n4, 20))
times = np.repeat(np.arange(10.0), n4 // 10)
ls, ls_time = 6.0, 1.5
cov = C.Matern32(ls, active_dims=slice(None, -1)) * C.ExpQuad(ls_time, active_dims=-1)
covo = O.Matern32(ls, active_dims=slice(None, -1)) * O.ExpQuad(ls_time, active_dims=-1)
Xt = np.concatenate([X4, times[:, None]], axis=1)
lm4 = np.ascontiguousarray(Xt[np.sort(np.random.default_rng(1).choice(n4, 1000, replace=False))])

def rs(a, b):
    d = np.asarray(a) - np.asarray(b)
    return f"rel_std {np.std(d) / np.std(b):.2e} max_abs/max {np.max(np.abs(d)) / np.max(np.abs(b)):.2e}"

TIGHT = dict(maxiter=20000, maxfun=100000, ftol=0.0, gtol=1e-9)
for tag, opts in (("default stop", None), ("converged", TIGHT)):
    old = dict(mb.inference.LBFGSB_OPTIONS)
    if opts:
        mb.inference.LBFGSB_OPTIONS.clear(); mb.inference.LBFGSB_OPTIONS.update(opts)
    est = mb.TimeSensitiveDensityEstimator(cov_func=cov, ls=ls, ls_time=ls_time, landmarks=lm4, check_rank=False)
    dens = est.fit_predict(X4, times)
    nn = np.asarray(est.nn_distances)
    mb.inference.LBFGSB_OPTIONS.clear(); mb.inference.LBFGSB_OPTIONS.update(old)
    ref = O.fit_density(Xt, cov_func=covo, landmarks=lm4, nn_distances=nn, d=20, ls=ls, lbfgsb_options=opts)
    perm = np.random.default_rng(5).permutation(1000)
    ref2 = O.fit_density(Xt, cov_func=covo, landmarks=np.ascontiguousarray(lm4[perm]), nn_distances=nn, d=20, ls=ls, lbfgsb_options=opts)
    print(f"[{tag}] oracle vs oracle(permuted landmarks): {rs(ref2.log_density_x, ref.log_density_x)}  nfev {ref.opt_state.num_fun_eval}/{ref2.opt_state.num_fun_eval}")
    print(f"[{tag}] cuda   vs oracle                    : {rs(dens, ref.log_density_x)}  nfev {est.opt_state.num_fun_eval}", flush=True)
