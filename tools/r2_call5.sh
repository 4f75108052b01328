#!/bin/bash
# Round 2, call 5: parity at BASELINE config shapes (new tests) + bench --config 2 on one GPU.
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 1500 python -m pytest tests/test_baseline_configs.py -m gpu -q -s > "$OUT/pytest_configs.txt" 2>&1
echo "config tests exit $?" >> "$OUT/pytest_configs.txt"; grep -E "^\[config|passed|failed|Error|assert" "$OUT/pytest_configs.txt" | head -40
timeout 900 python bench.py --config 2 --steps 3 --warmup 3 > "$OUT/bench_config2.json" 2> "$OUT/bench_config2.err"
echo "bench config 2 exit $?"; tail -3 "$OUT/bench_config2.err"
python - <<'P'
import json
d=json.load(open("gpurun_out/r2/bench_config2.json"))
print(d["ms_per_step"], d["value"], d["parity"], d["lbfgsb"])
P
