"""Kernel-level timings on one GPU (CUDA events on the library stream, best of 3 after a warm-up):

    python tools/bench_kernels.py [N=262144] [M=5000] [D=50] [k1|default]   (k1: the K1 lines only; default: K1 and
                                                                             the default K3 / K4 variant only)

K1 (fused distance + covariance build) on the int8 digit-slice kernel and on the FP64 DMMA kernel, K4 Gram and K3 TRSM
with the int8 slices on and off.  Prints ms, algorithmic GB/s (K1) and float64-equivalent TF/s (K3 / K4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mellon_b200 as mb
from mellon_b200 import cov as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
M = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 50
mb.setup_logging().setLevel("WARNING")
be = mb.get_backend()
rng = np.random.default_rng(0)
x = rng.random((N, D))
xu = np.ascontiguousarray(x[np.sort(rng.choice(N, M, replace=False))])
xd, xud = be.upload(x, sharded=True), be.upload(xu)
cov = C.Matern52(38.0)


def best(fn, reps=3):
    fn()
    be.sync()
    out = []
    for _ in range(reps):
        be.timer_start(2)
        fn()
        out.append(be.timer_stop(2))
    return min(out)


alg = 8.0 * (N * M + (N + M) * D)
for name, opt in (("int8 slices (tcgen05)", 1), ("FP64 DMMA", 0)):
    be.set_option("cov_i8", opt)
    be.prof_enable(True); be.prof_reset()
    ms = best(lambda: be.cov(cov, xd, xud, sharded=True))
    n_l, ms_k, _ = be.prof_read()["cov"]
    be.prof_enable(False)
    print(f"K1 {name:24s} N={N} M={M} D={D}: call {ms:8.3f} ms (kernel alone {ms_k / max(n_l, 1):8.3f} ms) = {alg / ms / 1e6:7.1f} GB/s "
          f"algorithmic incl. pack, {alg / (ms_k / max(n_l, 1)) / 1e6:7.1f} GB/s kernel only; scaled to N=1e6: {ms * 1e6 / N:6.2f} ms")
be.set_option("cov_i8", 1)
if len(sys.argv) > 4 and sys.argv[4] == "k1":
    sys.exit(0)

K = be.cov(cov, xd, xud, sharded=True)
Lp, info = be.cov_chol(cov, xu, 1e-6)
VARIANTS = (("int8 slices, A in TMEM", 1, 0, 0), ("int8 slices, 4 issuers", 1, 4, 0), ("int8 slices, 2 issuers", 1, 2, 0),
            ("int8 slices, 1 issuer", 1, 1, 0), ("FP64 DMMA", 0, 4, 0))
if len(sys.argv) > 4 and sys.argv[4] == "default":
    VARIANTS = (("int8, packs on side stream", 1, 4, 1), ("int8, one stream", 1, 4, 0), ("int8, packs on side stream", 1, 4, 1), ("int8, one stream", 1, 4, 0))
for name, opt, iss, ovl in VARIANTS:
    be.set_option("i8", opt)
    be.set_option("i8_issuers", iss)
    be.set_option("i8_overlap", ovl)
    ms_g = best(lambda: be.gram(K))
    Kc = be.copy(K)
    ms_t = best(lambda: be.trsm_right_lt(Lp, Kc), reps=2)
    del Kc
    print(f"K4 Gram {name:26s} N={N} r={M}: {ms_g:8.2f} ms = {N * M * M / ms_g / 1e9:6.1f} f64-equivalent TF/s; scaled to N=1e6: {ms_g * 1e6 / N:7.1f} ms")
    print(f"K3 TRSM {name:26s} N={N} m={M}: {ms_t:8.2f} ms = {N * M * M / ms_t / 1e9:6.1f} f64-equivalent TF/s; scaled to N=1e6: {ms_t * 1e6 / N:7.1f} ms")
be.set_option("i8", 1)
be.set_option("i8_issuers", 4)
be.set_option("i8_overlap", 0)
