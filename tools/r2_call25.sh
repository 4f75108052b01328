#!/bin/bash
# Round 2, call 25 (1 GPU): digit packs on a side stream under the int8 GEMMs ("i8_overlap").  Identical-bits test and
# kernel timings first; the whole GPU suite and the headline line only if those pass.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_kernels_parity.py -m gpu -q -p no:cacheprovider -k "side_stream or int8 or i8" > "$OUT/pytest_overlap.txt" 2>&1
rc=$?
echo "pytest overlap exit $rc"; tail -8 "$OUT/pytest_overlap.txt" | cut -c1-300
timeout 300 python tools/bench_kernels.py 524288 5000 50 default > "$OUT/bench_kernels_overlap.txt" 2>&1
echo "bench_kernels exit $?"; grep "^K[34]" "$OUT/bench_kernels_overlap.txt" | cut -c1-200
[ $rc -ne 0 ] && exit 1
timeout 600 python bench.py > "$OUT/bench_overlap_1gpu.json" 2> "$OUT/bench_overlap_1gpu.err"
echo "bench exit $?"; tail -3 "$OUT/bench_overlap_1gpu.err" | cut -c1-300
python - <<P
import json
try:
    d=json.loads([l for l in open("$OUT/bench_overlap_1gpu.json") if l.startswith("{")][-1])
    print(d["n_gpus"], d["metric"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["lbfgsb"], d["log_density_sha256"])
    p=d["parity"]; print({k:p.get(k) for k in ("rel_std_err_log_density","ok","nfev_cuda","nfev_cpu")})
    for k,v in d["kernels"].items(): print("  ", k[:40], v)
    print(d["clocks"])
except Exception as e: print("no line", e)
P
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu_overlap.txt" 2>&1
echo "pytest exit $?"; tail -8 "$OUT/pytest_gpu_overlap.txt" | cut -c1-300
