#!/bin/bash
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_kernels_parity.py -m gpu -q -x -k "k1_ or cov_build or active_dims or algebra or mutable" > "$OUT/pytest_k1.txt" 2>&1
echo "k1 tests exit $?" >> "$OUT/pytest_k1.txt"; tail -6 "$OUT/pytest_k1.txt" | cut -c1-300
timeout 600 python tools/bench_kernels.py 262144 5000 50 > "$OUT/bench_kernels_c.txt" 2>&1; head -2 "$OUT/bench_kernels_c.txt" | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cov_i8_kernel -s 1 -c 1 -f -o "$OUT/ncu_cov_i8_r02" python tools/bench_kernels.py 262144 5000 50 > "$OUT/ncu_cov_i8.log" 2>&1
echo "ncu exit $?"
