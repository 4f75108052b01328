#!/bin/bash
# Round 2, call 2: GPU suite + headline bench on one GPU after the reduction-tree refactor.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu_a.txt" 2>&1
echo "gpu suite exit $?" >> "$OUT/pytest_gpu_a.txt"
tail -15 "$OUT/pytest_gpu_a.txt"
timeout 900 python bench.py --steps 3 --warmup 3 > "$OUT/bench_a.json" 2> "$OUT/bench_a.err"
echo "bench exit $?"
tail -c 3000 "$OUT/bench_a.json"; tail -5 "$OUT/bench_a.err"
