#!/bin/bash
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_kernels_parity.py -m gpu -q -x -k "issue_variants or k1_" > "$OUT/pytest_atmem.txt" 2>&1
echo "tests exit $?" >> "$OUT/pytest_atmem.txt"; tail -12 "$OUT/pytest_atmem.txt" | cut -c1-300
timeout 600 python tools/bench_kernels.py 262144 5000 50 > "$OUT/bench_kernels_e.txt" 2>&1; cat "$OUT/bench_kernels_e.txt" | cut -c1-300
