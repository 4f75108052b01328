#!/bin/bash
# Round 2, call 3: whole GPU suite (no -x) with the int8 Gram in the product + headline bench.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_kernels_parity.py -m gpu -q -k "int8" > "$OUT/pytest_i8.txt" 2>&1
echo "i8 tests exit $?" >> "$OUT/pytest_i8.txt"; tail -25 "$OUT/pytest_i8.txt"
timeout 1200 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu_b.txt" 2>&1
echo "gpu suite exit $?" >> "$OUT/pytest_gpu_b.txt"
tail -15 "$OUT/pytest_gpu_b.txt"
timeout 900 python bench.py --steps 3 --warmup 3 > "$OUT/bench_b.json" 2> "$OUT/bench_b.err"
echo "bench exit $?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2/bench_b.json"))
print(d["ms_per_step"], d["value"], d["parity"]["rel_std_err_log_density"], d["lbfgsb"])
for k,v in d["kernels"].items(): print(k, v)
P
tail -5 "$OUT/bench_b.err"
