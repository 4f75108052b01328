#!/bin/bash
# Round 2, call 26 (1 GPU): the final build ("i8_overlap" off by default): headline line, then the whole GPU suite.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 600 python bench.py > "$OUT/bench_final2_1gpu.json" 2> "$OUT/bench_final2_1gpu.err"
echo "bench exit $?"; tail -3 "$OUT/bench_final2_1gpu.err" | cut -c1-300
python - <<P
import json
try:
    d=json.loads([l for l in open("$OUT/bench_final2_1gpu.json") if l.startswith("{")][-1])
    print(d["n_gpus"], d["metric"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["lbfgsb"], d["log_density_sha256"])
    p=d["parity"]; print({k:p.get(k) for k in ("rel_std_err_log_density","ok","nfev_cuda","nfev_cpu")})
    for k,v in d["kernels"].items(): print("  ", k[:40], v)
    print(d["clocks"])
except Exception as e: print("no line", e)
P
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu_final2.txt" 2>&1
echo "pytest exit $?"; tail -8 "$OUT/pytest_gpu_final2.txt" | cut -c1-300
