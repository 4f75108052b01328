#!/bin/bash
# Round 2, call 9 (4 GPUs): invariance check, headline, configs 4 and 3, and a host-clock stage trace of one step.
set -u
bash tools/r2_call8.sh 4 "headline 4 3"
OUT=gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tools/trace_step.py > "$OUT/trace_step_4gpu.txt" 2>&1
grep -A 12 "rep 1" "$OUT/trace_step_4gpu.txt"
