"""Small-shape pass through every hand-written kernel family, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_subset.py
    compute-sanitizer --tool racecheck python tools/sanitize_subset.py

K1 (FP64 DMMA kernel with its bulk-TMA + mbarrier ring, the register-tile kernel, and the tcgen05 int8 digit-slice
kernel), K5 / K6 (bulk-TMA ring streaming kernel and the register-fused one), K2 Cholesky (128-wide leaf + updates),
K3 TRSM and K4 Gram (FP64 DMMA tiles and the int8 digit-slice GEMMs), K7.  Each result is checked against NumPy so a
silent wrong answer under the tool is caught too."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.linalg import solve_triangular

import mellon_b200 as mb
from mellon_b200 import cov as C
from oracle import mellon_oracle as O

mb.setup_logging().setLevel("WARNING")
be = mb.get_backend()
rng = np.random.default_rng(0)


def check(name, a, b, tol):
    err = float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(b))))
    print(f"{name:44s} rel err {err:.2e}")
    assert err < tol, name


x, y = rng.standard_normal((700, 20)), rng.standard_normal((150, 20))
ref = O.Matern52(1.7)(x, y)
for opt, nm in ((0, "K1 DMMA (bulk TMA + mbarrier)"), (1, "K1 DFMA register tiles")):
    be.set_option("cov", opt)
    be.set_option("cov_i8", 0)
    check(nm, be.cov(C.Matern52(1.7), x, y).numpy(), ref, 1e-12)
be.set_option("cov", 0)
be.set_option("cov_i8", 2)
check("K1 tcgen05 int8 digit slices", be.cov(C.Matern52(1.7), x, y).numpy(), ref, 1e-12)
be.set_option("cov_i8", 1)
w = rng.standard_normal(150)
check("K7 predict_mean", be.predict_mean(C.Matern52(1.7), x, y, w, 0.5), 0.5 + ref @ w, 1e-12)

L = rng.standard_normal((900, 128)) / 11.0
nn = rng.random(900) * 0.5 + 0.05
z = rng.standard_normal(128) * 0.3
V, Vdr = O.nn_constants(nn, 7.0)
for opt, nm in ((0, "K5 / K6 bulk-TMA ring"), (2, "K5 / K6 register-fused"), (1, "K5 / K6 two-pass")):
    be.set_option("lossgrad", opt)
    st = be.objective(L, V, float(np.sum(Vdr)), -3.0, 128)
    loss, grad = be.loss_grad(st, z)
    lref, gref = O.loss_and_grad(L, nn, 7.0, -3.0, z, 128)
    check(nm + " gradient", grad, gref, 1e-11)
    check(nm + " Hessian diagonal", be.hess_diag(st, z), O.hessian_diag(L, nn, 7.0, -3.0, z), 1e-11)
be.set_option("lossgrad", 0)

A = rng.standard_normal((300, 300))
S = A @ A.T + 300 * np.eye(300)
Sd = be.upload(S.copy())
assert be.potrf(Sd) == 0
Lp = np.linalg.cholesky(S)
check("K2 Cholesky", np.tril(Sd.numpy()), Lp, 1e-12)
X = rng.standard_normal((600, 300))
for opt, nm in ((0, "FP64 DMMA"), (2, "tcgen05 int8 slices")):
    be.set_option("i8", opt)
    check("K3 TRSM " + nm, be.trsm_right_lt(be.upload(Lp), be.upload(X.copy(), sharded=True)).numpy(),
          solve_triangular(Lp, X.T, lower=True).T, 1e-11)
    Ld = be.upload(L, sharded=True)
    check("K4 Gram " + nm, be.gram(Ld).numpy(), L.T @ L, 1e-12)
    check("ridge init " + nm, be.ridge_init(Ld, nn), O.ridge_normal_equations(L, nn), 1e-9)
be.set_option("i8", 1)
print("SANITIZE_SUBSET done")
