#!/bin/bash
# Round 2, call 10 (one GPU): cheaper MMA issue path -> tests, bench, launch list and ncu --set full of the int8 kernels.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu_d.txt" 2>&1
echo "gpu suite exit $?" >> "$OUT/pytest_gpu_d.txt"; tail -6 "$OUT/pytest_gpu_d.txt"
timeout 900 python bench.py --steps 3 --warmup 3 > "$OUT/bench_d.json" 2> "$OUT/bench_d.err"
echo "bench exit $?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2/bench_d.json"))
print(d["ms_per_step"], d["value"], d["parity"]["rel_std_err_log_density"], d["lbfgsb"])
for k,v in d["kernels"].items(): print(k, v)
P
SMALL="--cells 200000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --predict-queries 0 --no-clocks"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_r02.csv" python bench.py $SMALL > "$OUT/launches_r02.log" 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_i8_kernel -s 2 -c 1 -f -o "$OUT/ncu_gram_i8_r02" python bench.py $SMALL > "$OUT/ncu_gram.log" 2>&1
echo "ncu gram exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_i8_kernel -s 0 -c 1 -f -o "$OUT/ncu_gemm_nt_i8_r02" python bench.py $SMALL > "$OUT/ncu_nt.log" 2>&1
echo "ncu nt exit $?"
