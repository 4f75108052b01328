"""Parity at BASELINE.json's configurations (shapes the CPU oracle can finish in minutes), one GPU:

    python tools/parity_configs.py > profiles/parity_configs_rNN.txt

For each configuration the identical inputs (cells, landmarks, exact nn-distances) go through the CUDA path
and through the CPU oracle; reported: the reference's own acceptance metric std(a - b) / std(b)
(tests/test_density_estimator.py:30-44), max |a - b| / max |b|, and the L-BFGS-B evaluation counts."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mellon_b200 as mb
from mellon_b200 import cov as C
from oracle import mellon_oracle as O

mb.setup_logging().setLevel("WARNING")
be = mb.get_backend()


def report(name, a, b, extra=""):
    a, b = np.asarray(a), np.asarray(b)
    d = a - b
    print(f"{name:58s} rel_std_err {np.std(d) / np.std(b):.3e}   max_abs/max_abs {np.max(np.abs(d)) / np.max(np.abs(b)):.3e} {extra}",
          flush=True)


def landmarks_of(x, m, seed=1):
    return np.ascontiguousarray(x[np.sort(np.random.default_rng(seed).choice(x.shape[0], m, replace=False))])


# config 1: README smoke test, defaults -> FULL GP
X = np.random.default_rng(0).random((100, 10))
est = mb.DensityEstimator()
dens = est.fit_predict(X)
ref = O.fit_density(X)
report("config 1  rand(100,10) defaults (FULL, Matern52)", dens, ref.log_density_x)
Y = np.random.default_rng(1).random((100, 10))
report("config 1  predict(rand(100,10))", est.predict(Y), O.predict_density(ref, X, Y))

# config 2: N = 100k, D = 50, M = 5000, ExpQuad, sparse Cholesky
n, m = int(os.environ.get("PARITY_N2", 100_000)), 5000
X = np.random.default_rng(0).random((n, 50))
lm = landmarks_of(X, m)
nn = be.nn_distances(X)
t0 = time.perf_counter()
est = mb.DensityEstimator(cov_func_curry=C.ExpQuad, landmarks=lm, nn_distances=nn, check_rank=False)
dens = est.fit_predict(X)
t_gpu = time.perf_counter() - t0
t0 = time.perf_counter()
ref = O.fit_density(X, cov_func_curry=O.ExpQuad, landmarks=lm, nn_distances=nn)
t_cpu = time.perf_counter() - t0
report(f"config 2  N={n} D=50 M={m} ExpQuad sparse_cholesky", dens, ref.log_density_x,
       f"nfev gpu/cpu {est.opt_state.num_fun_eval}/{getattr(ref.opt_state, 'num_fun_eval', getattr(ref.opt_state, 'nfev', '?'))}  wall gpu {t_gpu:.2f} s cpu {t_cpu:.1f} s")
# config 5 shape: out-of-sample predict on the fitted model
Y = np.random.default_rng(2).random((200_000, 50))
report("config 5  predict 200k queries on the config-2 model", est.predict(Y), O.predict_density(ref, X, Y))
del est

# config 3 shape: Matern52, rank = 2000 (int -> sparse Nystroem), N reduced so that the oracle's QR + eigh finish
n3 = int(os.environ.get("PARITY_N3", 30_000))
X3 = np.ascontiguousarray(X[:n3])
nn3 = be.nn_distances(X3)
est = mb.DensityEstimator(landmarks=lm, nn_distances=nn3, rank=2000, check_rank=False)
dens = est.fit_predict(X3)
ref = O.fit_density(X3, landmarks=lm, nn_distances=nn3, rank=2000)
report(f"config 3  N={n3} D=50 M={m} Matern52 Nystroem rank=2000", dens, ref.log_density_x)
del est

# config 4 shape: time-sensitive, Matern32 (space) x ExpQuad (time), 10 time points
n4 = int(os.environ.get("PARITY_N4", 20_000))
X4 = np.random.default_rng(0).random((n4, 20))
times = np.repeat(np.arange(10.0), n4 // 10)
ls, ls_time = 6.0, 1.5
cov = C.Matern32(ls, active_dims=slice(None, -1)) * C.ExpQuad(ls_time, active_dims=-1)
covo = O.Matern32(ls, active_dims=slice(None, -1)) * O.ExpQuad(ls_time, active_dims=-1)
Xt = np.concatenate([X4, times[:, None]], axis=1)
lm4 = landmarks_of(Xt, 1000)
est = mb.TimeSensitiveDensityEstimator(cov_func=cov, ls=ls, ls_time=ls_time, landmarks=lm4, check_rank=False)
dens = est.fit_predict(X4, times)
ref = O.fit_density(Xt, cov_func=covo, landmarks=lm4, nn_distances=np.asarray(est.nn_distances), d=20, ls=ls)
report(f"config 4  N={n4} D=20+time M=1000 Matern32 x ExpQuad(time)", dens, ref.log_density_x)
print("PARITY_CONFIGS done")
