"""Launch each hot kernel of the path exactly once on config-shaped operands, in a fixed order, so that
one `ncu -k regex:"cov_tile_kernel|cov_mma_kernel|gemm_dmma_kernel|stream_rows_kernel" -c 6` capture holds:
K1 (K_NM build), K4 (Gram SYRK, k-major operands), a K3-shaped update GEMM, K5 (loss+grad, twice), K6.

    python tools/profile_kernels.py [N] [M] [D]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mellon_b200 as mb
from mellon_b200 import cov as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 50
be = mb.get_backend()
rng = np.random.default_rng(0)
x = rng.random((N, D))
lm = np.ascontiguousarray(x[:M])
cov = C.Matern52(ls=38.0)
xd = be.upload(x, sharded=True)
be.sync()
K = be.cov(cov, xd, lm, sharded=True)          # K1
be.sync()
G = be.gram(K)                                  # K4: SYRK over the cell axis
be.sync()
W = be.upload(rng.standard_normal((1280, M)))
out = be.gemm(K, W, trans_b=True)               # K3-shaped update: (N x M) . (1280 x M)^T
be.sync()
V = rng.random(N) * 0.1 - 20.0
st = be.objective(K, V, 0.0, -10.0, M)
z = rng.standard_normal(M) * 1e-3
for _ in range(2):
    loss, grad = be.loss_grad(st, z)            # K5
h = be.hess_diag(st, z)                         # K6
be.sync()
print("ok", float(loss), float(np.abs(grad).max()), float(h.max()), be.launch_count())
