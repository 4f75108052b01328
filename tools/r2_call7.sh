#!/bin/bash
# Round 2, call 7 (one GPU): functional run of bench --config 3 / 4 / 5 at reduced sizes before the multi-GPU calls.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
for cfg in "3 --cells 200000" "4 --cells 100000" "5 --predict-queries 1000000"; do
  tag=$(echo $cfg | cut -d' ' -f1)
  timeout 900 python bench.py --config $cfg --steps 2 --warmup 1 > "$OUT/bench_small_config$tag.json" 2> "$OUT/bench_small_config$tag.err"
  echo "config $tag exit $?"; tail -3 "$OUT/bench_small_config$tag.err" | cut -c1-800
  python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench_small_config$tag.json") if l.startswith("{")][-1])
    print(d["config"]["workload"]); print(d["metric"], d["ms_per_step"], d["value"], d["lbfgsb"]); print(d["parity"])
    for k,v in d["kernels"].items(): print("  ", k, v)
except Exception as e: print("no line", e)
P
done
