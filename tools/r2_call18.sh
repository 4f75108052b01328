#!/bin/bash
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_kernels_parity.py tests/test_kmeans.py -m gpu -q -x -k "nn_ or kmeans or k_means or seed_rows" > "$OUT/pytest_nn.txt" 2>&1
echo "nn tests exit $?" >> "$OUT/pytest_nn.txt"; tail -15 "$OUT/pytest_nn.txt" | cut -c1-300
python - <<'P'
import time, numpy as np
import mellon_b200 as mb
mb.setup_logging().setLevel("WARNING")
be = mb.get_backend()
x = np.random.default_rng(0).random((1_000_000, 50))
for opt, name in ((1, "tcgen05 int8 digit slices"), (0, "FP64 DFMA register tiles")):
    be.set_option("cov_i8", opt)
    be.nn_distances(x[:20000])
    t = time.perf_counter(); d, i = be.nn_distances(x, return_index=True); dt = time.perf_counter() - t
    print(f"exact 1-NN, N = 1e6, D = 50, {name}: {dt:.2f} s (host in, host out), checksum {d.sum():.10f} {int(i.sum())}")
be.set_option("cov_i8", 1)
P
