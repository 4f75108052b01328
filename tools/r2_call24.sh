#!/bin/bash
# Round 2, call 24 (1 GPU): digit packing by the byte trick (mb_i8.cu fixed54x16 / digits_of4).  Kernel timings, the
# headline line, the whole GPU suite, and a fresh launch list of the final build.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 300 python tools/bench_kernels.py 262144 5000 50 default > "$OUT/bench_kernels_pack.txt" 2>&1
echo "bench_kernels exit $?"; cat "$OUT/bench_kernels_pack.txt" | cut -c1-200
timeout 600 python bench.py > "$OUT/bench_pack_1gpu.json" 2> "$OUT/bench_pack_1gpu.err"
echo "bench exit $?"; tail -3 "$OUT/bench_pack_1gpu.err" | cut -c1-300
python - <<P
import json
try:
    d=json.loads([l for l in open("$OUT/bench_pack_1gpu.json") if l.startswith("{")][-1])
    print(d["n_gpus"], d["metric"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["lbfgsb"], d["log_density_sha256"])
    p=d["parity"]; print({k:p.get(k) for k in ("rel_std_err_log_density","ok","nfev_cuda","nfev_cpu")})
    for k,v in d["kernels"].items(): print("  ", k[:40], v)
    print(d["roofline"]); print(d["clocks"])
except Exception as e: print("no line", e)
P
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu_pack.txt" 2>&1
echo "pytest exit $?"; tail -15 "$OUT/pytest_gpu_pack.txt" | cut -c1-300
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/launches_pack.csv" \
  python bench.py --cells 250000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --predict-queries 0 --no-clocks > "$OUT/launches_pack.log" 2>&1
echo "ncu exit $?"; tail -2 "$OUT/launches_pack.log" | cut -c1-200
