#!/bin/bash
# Round 2, call 4: TRSM updates on int8 digit slices.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_kernels_parity.py -m gpu -q -k "int8 or trsm" > "$OUT/pytest_i8b.txt" 2>&1
echo "i8 tests exit $?" >> "$OUT/pytest_i8b.txt"; tail -25 "$OUT/pytest_i8b.txt"
timeout 900 python bench.py --steps 3 --warmup 3 > "$OUT/bench_c.json" 2> "$OUT/bench_c.err"
echo "bench exit $?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2/bench_c.json"))
print(d["ms_per_step"], d["value"], d["parity"]["rel_std_err_log_density"], d["lbfgsb"])
for k,v in d["kernels"].items(): print(k, v)
P
tail -5 "$OUT/bench_c.err"
timeout 900 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu_c.txt" 2>&1
echo "gpu suite exit $?" >> "$OUT/pytest_gpu_c.txt"; tail -8 "$OUT/pytest_gpu_c.txt"
