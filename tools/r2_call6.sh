#!/bin/bash
# Round 2, call 6 (N GPUs): world-size invariance + headline bench on N ranks.   usage: r2_call6.sh N
set -u
N=${1:-2}
OUT=gpurun_out/r2
mkdir -p "$OUT"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29511 tools/check_multi_gpu.py > "$OUT/check_multi_gpu_$N.txt" 2>&1
echo "check exit $?"; grep -E "world=|MULTI_GPU|Error|error" "$OUT/check_multi_gpu_$N.txt" | cut -c1-1500 | head -20
timeout 900 $RUN --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > "$OUT/bench_${N}gpu.json" 2> "$OUT/bench_${N}gpu.err"
echo "bench exit $?"; tail -3 "$OUT/bench_${N}gpu.err" | cut -c1-600
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r2/bench_${N}gpu.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["lbfgsb"], d["log_density_sha256"])
print(d["parity"])
for k,v in d["kernels"].items(): print(k, v)
P
