"""Device timings (CUDA events through the C ABI) of the M x M linear algebra of one fit: Cholesky, vector
triangular solves, and the N x M TRSM / Gram at a reduced N.   python tools/time_linalg.py [M] [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mellon_b200 as mb
from mellon_b200 import cov as C

M = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
be = mb.get_backend()
rng = np.random.default_rng(0)
x = rng.random((N, 50))
lm = np.ascontiguousarray(x[:M])
cov = C.Matern52(ls=38.0)

def timed(name, fn, reps=3):
    best = 1e30
    for _ in range(reps):
        be.sync(); be.timer_start(2); out = fn(); ms = be.timer_stop(2); best = min(best, ms)
    print(f"{name:34s} {best:9.3f} ms", flush=True)
    return out

for variant, tag in ((0, "leaf128 + inverse GEMM panels"), (1, "leaf32 substitution")):
    be.set_option("trsm", variant)
    def chol():
        Lp, info = be.cov_chol(cov, lm, 1e-6)
        assert info == 0
        return Lp
    Lp = timed(f"cov_chol M={M} [{tag}]", chol)
    b = be.upload(rng.standard_normal(M))
    timed(f"tri_solve fwd [{tag}]", lambda: be.tri_solve_dev(Lp, b, trans=False))
    timed(f"tri_solve bwd [{tag}]", lambda: be.tri_solve_dev(Lp, b, trans=True))
be.set_option("trsm", 0)
xd = be.upload(x, sharded=True)
K = timed(f"K1 cov N={N}", lambda: be.cov(cov, xd, lm, sharded=True), reps=2)
timed(f"K3 trsm N={N}", lambda: be.trsm_right_lt(Lp, K), reps=2)
timed(f"K4 gram N={N}", lambda: be.gram(K), reps=2)
