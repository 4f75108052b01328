#!/bin/bash
# Round 2, call 8 (N GPUs): invariance check, headline bench and the BASELINE configs quoted on N GPUs.  usage: r2_call8.sh N "cfg cfg ..."
set -u
N=${1:-8}
CFGS=${2:-"headline 3 5"}
OUT=gpurun_out/r2
mkdir -p "$OUT"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 tools/check_multi_gpu.py > "$OUT/check_multi_gpu_$N.txt" 2>&1
echo "check exit $?"; grep -E "world=|MULTI_GPU|Error|error" "$OUT/check_multi_gpu_$N.txt" | cut -c1-1200 | head -20
port=29520
for cfg in $CFGS; do
  port=$((port+1))
  timeout 900 $RUN --master-port $port bench.py --config $cfg --gpus $N --steps 3 --warmup 3 > "$OUT/bench_${N}gpu_$cfg.json" 2> "$OUT/bench_${N}gpu_$cfg.err"
  echo "bench $cfg exit $?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" "$OUT/bench_${N}gpu_$cfg.err" | tail -3 | cut -c1-600
  python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench_${N}gpu_$cfg.json") if l.startswith("{")][-1])
    print(d["config"]["workload"]); print(d["n_gpus"], d["metric"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["lbfgsb"])
    print(d["parity"])
    for k,v in d["kernels"].items(): print("  ", k, v)
    print("  predict", d.get("predict"))
except Exception as e: print("no line", e)
P
done
