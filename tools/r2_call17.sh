#!/bin/bash
# Round 2, call 17 (one GPU): final evidence — suite, config parity lines, bench, launch list, ncu --set full of the hot kernels.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu_final.txt" 2>&1
echo "gpu suite exit $?" >> "$OUT/pytest_gpu_final.txt"; tail -4 "$OUT/pytest_gpu_final.txt" | cut -c1-200
timeout 900 python -m pytest tests/test_baseline_configs.py -m gpu -q -s 2>&1 | grep -E "^\[config|passed|failed" > "$OUT/parity_configs_lines.txt"; cat "$OUT/parity_configs_lines.txt" | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > "$OUT/bench_final_1gpu.json" 2> "$OUT/bench_final_1gpu.err"
echo "bench exit $?"; tail -2 "$OUT/bench_final_1gpu.err" | cut -c1-300
timeout 600 python bench.py --config 2 --steps 3 --warmup 3 > "$OUT/bench_final_config2.json" 2> "$OUT/bench_final_config2.err"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_final_reference.json" 2> "$OUT/bench_final_reference.err"
echo "reference exit $?"
SMALL="--cells 250000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --predict-queries 0 --no-clocks"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$OUT/launches_r02_final.csv" python bench.py $SMALL > "$OUT/launches_final.log" 2>&1
echo "launch list exit $?"
for k in cov_i8_kernel gram_i8_kernel gemm_nt_i8_kernel stream_rows_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o "$OUT/ncu_final_$k" python bench.py $SMALL > "$OUT/ncu_final_$k.log" 2>&1
  echo "ncu $k exit $?"
done
