#!/bin/bash
# Round 2, final multi-GPU lines: headline (+ optional configs) on N ranks.  usage: r2_call23.sh N "cfg cfg ..." [check]
set -u
N=${1:-8}
CFGS=${2:-"headline 3"}
OUT=gpurun_out/r2
mkdir -p "$OUT"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${3:-}" = "check" ]; then
  timeout 600 $RUN --master-port 29511 tools/check_multi_gpu.py > "$OUT/check_multi_gpu_final_$N.txt" 2>&1
  echo "check exit $?"; grep -E "world=|MULTI_GPU" "$OUT/check_multi_gpu_final_$N.txt" | cut -c1-900
fi
port=29540
for cfg in $CFGS; do
  port=$((port+1))
  timeout 900 $RUN --master-port $port bench.py --config $cfg --gpus $N --steps 3 --warmup 3 > "$OUT/bench_final_${N}gpu_$cfg.json" 2> "$OUT/bench_final_${N}gpu_$cfg.err"
  echo "bench $cfg exit $?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" "$OUT/bench_final_${N}gpu_$cfg.err" | tail -3 | cut -c1-400
  python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench_final_${N}gpu_$cfg.json") if l.startswith("{")][-1])
    print(d["config"]["workload"]); print(d["n_gpus"], d["metric"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["lbfgsb"])
    p=d["parity"]; print({k:p.get(k) for k in ("rel_std_err_log_density","ok","oracle_vs_oracle_floor","vs_one_gpu")})
    for k,v in d["kernels"].items(): print("  ", k[:40], v)
except Exception as e: print("no line", e)
P
done
