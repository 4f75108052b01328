#!/bin/bash
# Round 2, call 27 (1 GPU): the whole GPU suite on the final library (per-device kernel configuration flags).
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
timeout 240 python -m pytest tests -m gpu -q -x -p no:cacheprovider > "$OUT/pytest_gpu_final3.txt" 2>&1
echo "pytest exit $?"; tail -6 "$OUT/pytest_gpu_final3.txt" | cut -c1-300
