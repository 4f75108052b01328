#!/bin/bash
# Round 2, call 14 (one GPU): full GPU suite, kernel timings, headline bench, sanitizer subset.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu_e.txt" 2>&1
echo "gpu suite exit $?" >> "$OUT/pytest_gpu_e.txt"; tail -6 "$OUT/pytest_gpu_e.txt" | cut -c1-300
timeout 600 python tools/bench_kernels.py 262144 5000 50 > "$OUT/bench_kernels_d.txt" 2>&1; cat "$OUT/bench_kernels_d.txt" | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > "$OUT/bench_g.json" 2> "$OUT/bench_g.err"
echo "bench exit $?"; tail -3 "$OUT/bench_g.err" | cut -c1-300
python - <<'P'
import json
d=json.load(open("gpurun_out/r2/bench_g.json"))
print(d["ms_per_step"], d["value"], d["parity"]["rel_std_err_log_density"], d["lbfgsb"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"])
for k,v in d["kernels"].items(): print(k, v)
P
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_subset.py > "$OUT/sanitizer_memcheck.txt" 2>&1
echo "memcheck exit $?"; tail -4 "$OUT/sanitizer_memcheck.txt" | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_subset.py > "$OUT/sanitizer_racecheck.txt" 2>&1
echo "racecheck exit $?"; tail -4 "$OUT/sanitizer_racecheck.txt" | cut -c1-200
