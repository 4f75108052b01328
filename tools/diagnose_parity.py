"""Where does a log-density difference at config-2 scale come from?  Same inputs through the CPU oracle, the
oracle with a rounding-level reformulation (landmarks permuted: identical mathematics, different rounding), and
the CUDA path under each kernel option."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mellon_b200 as mb
from mellon_b200 import cov as C
from oracle import mellon_oracle as O

mb.setup_logging().setLevel("WARNING")
be = mb.get_backend()
n = int(os.environ.get("DIAG_N", 30000)); m = int(os.environ.get("DIAG_M", 5000))
covname = os.environ.get("DIAG_COV", "ExpQuad")
X = np.random.default_rng(0).random((n, 50))
lm = np.ascontiguousarray(X[np.sort(np.random.default_rng(1).choice(n, m, replace=False))])
nn = be.nn_distances(X)

def rs(a, b):
    d = np.asarray(a) - np.asarray(b)
    return f"rel_std {np.std(d) / np.std(b):.2e} max_abs/max {np.max(np.abs(d)) / np.max(np.abs(b)):.2e}"

ref = O.fit_density(X, cov_func_curry=getattr(O, covname), landmarks=lm, nn_distances=nn)
print(f"{covname} N={n} M={m}: oracle nfev {ref.opt_state.num_fun_eval} loss {ref.loss:.10e}", flush=True)
perm = np.random.default_rng(5).permutation(m)
ref2 = O.fit_density(X, cov_func_curry=getattr(O, covname), landmarks=np.ascontiguousarray(lm[perm]), nn_distances=nn)
print("oracle vs oracle(landmarks permuted)      ", rs(ref2.log_density_x, ref.log_density_x), "nfev", ref2.opt_state.num_fun_eval, flush=True)
tight = dict(maxiter=20000, maxfun=100000, ftol=0.0, gtol=1e-9)
for opts in ({}, {"trsm": 1}, {"cov": 1}, {"gemm": 3}, {"trsm": 1, "cov": 1, "gemm": 3}):
    for k, v in opts.items():
        be.set_option(k, v)
    est = mb.DensityEstimator(cov_func_curry=getattr(C, covname), landmarks=lm, nn_distances=nn, check_rank=False)
    dens = est.fit_predict(X)
    # stage-wise: the device factor against the oracle's
    L = np.asarray(est.L)
    dL = np.max(np.abs(L - ref.L)) / np.max(np.abs(ref.L))
    z0 = np.asarray(est.initial_value)
    dz0 = np.max(np.abs(z0 - ref.initial_value)) / np.max(np.abs(ref.initial_value))
    print(f"cuda {str(opts):38s}", rs(dens, ref.log_density_x), f"nfev {est.opt_state.num_fun_eval} loss {est.losses[-1]:.10e} |dL| {dL:.2e} |dz0| {dz0:.2e}", flush=True)
    for k in opts:
        be.set_option(k, 0)
    del est
