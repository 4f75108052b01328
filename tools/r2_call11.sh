#!/bin/bash
# Round 2, call 11 (one GPU): single-issuer round-robin MMA order + K1 on int8 digit slices.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_kernels_parity.py -m gpu -q -x -k "int8 or trsm or gram or cov_build or active_dims or algebra" > "$OUT/pytest_i8d.txt" 2>&1
echo "i8 tests exit $?" >> "$OUT/pytest_i8d.txt"; tail -25 "$OUT/pytest_i8d.txt" | cut -c1-300
timeout 600 python tools/bench_kernels.py > "$OUT/bench_kernels_b.txt" 2>&1; cat "$OUT/bench_kernels_b.txt" | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > "$OUT/bench_f.json" 2> "$OUT/bench_f.err"
echo "bench exit $?"; tail -3 "$OUT/bench_f.err" | cut -c1-300
python - <<'P'
import json
d=json.load(open("gpurun_out/r2/bench_f.json"))
print(d["ms_per_step"], d["value"], d["parity"]["rel_std_err_log_density"], d["lbfgsb"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
for k,v in d["kernels"].items(): print(k, v)
P
