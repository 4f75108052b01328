#!/bin/bash
# First gpurun call of round 2: confirm on hardware what round 1 finished after its GPU minutes were spent, then run the
# int8 digit-slice prototypes (each under its own `timeout`: every mbarrier wait in them is bounded, this is the second fence).
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'      (about 15-20 GPU-minutes when nothing hangs)
#
# Everything lands in gpurun_out/round2_first/ ; nothing here is a bench value (see bench.py for those).
set -u
OUT=gpurun_out/round2_first
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1

# 1. the tests written without a GPU (everything marked run_last: FunctionEstimator, gemm beta, row scaling / diagonal vector,
#    the estimator tests that now take their nn_distances from the device)
timeout 900 python -m pytest tests -m "gpu and run_last" -q > "$OUT/pytest_new.txt" 2>&1
echo "new tests exit $?" >> "$OUT/pytest_new.txt"

# 2. prototypes: Gram matrix and TRSM on tcgen05 int8 digit slices (small first: correctness; then a timing size)
NVCC="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo"
(cd tools && $NVCC -o gram_i8_proto gram_i8_proto.cu && $NVCC -DDIGIT_BITS=7 -o gram_i8_proto7 gram_i8_proto.cu && $NVCC -DCLUSTER=2 -o gram_i8_proto_c2 gram_i8_proto.cu) > "$OUT/build.txt" 2>&1
for args in "4096 256" "65536 640" "131072 5000"; do
  echo "== gram $args" >> "$OUT/gram_i8_proto.txt"
  timeout 120 tools/gram_i8_proto $args >> "$OUT/gram_i8_proto.txt" 2>&1
  echo "exit $?" >> "$OUT/gram_i8_proto.txt"
done
echo "== gram 65536 640, 7-bit digits (36 products, five issuing warps)" >> "$OUT/gram_i8_proto.txt"
timeout 120 tools/gram_i8_proto7 65536 640 >> "$OUT/gram_i8_proto.txt" 2>&1
echo "exit $?" >> "$OUT/gram_i8_proto.txt"
for args in "4096 256" "131072 5000"; do     # only after the plain build is exact: A block multicast across a 2-CTA cluster
  echo "== gram $args, CLUSTER=2" >> "$OUT/gram_i8_proto.txt"
  timeout 120 tools/gram_i8_proto_c2 $args >> "$OUT/gram_i8_proto.txt" 2>&1
  echo "exit $?" >> "$OUT/gram_i8_proto.txt"
done
for args in "1024 256" "8192 640" "131072 2560"; do
  echo "== trsm $args" >> "$OUT/gram_i8_proto.txt"
  timeout 180 tools/gram_i8_proto trsm $args >> "$OUT/gram_i8_proto.txt" 2>&1
  echo "exit $?" >> "$OUT/gram_i8_proto.txt"
done

# 2b. K1 on int8 slices, v3 (four issuing warps, drain-then-compute epilogue) next to v2 for the before / after
(cd tools && $NVCC -o k1_i8_proto_v3 k1_i8_proto_v3.cu && $NVCC -o k1_i8_proto_v2 k1_i8_proto_v2.cu) >> "$OUT/build.txt" 2>&1
for exe in k1_i8_proto_v3 k1_i8_proto_v2; do
  echo "== $exe 131072 5120" >> "$OUT/k1_i8_proto.txt"
  timeout 120 tools/$exe 131072 5120 >> "$OUT/k1_i8_proto.txt" 2>&1
  echo "exit $?" >> "$OUT/k1_i8_proto.txt"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_i8_kernel -c 1 -o "$OUT/k1_i8_proto_v3" \
  tools/k1_i8_proto_v3 131072 5120 > "$OUT/ncu_k1.txt" 2>&1

# 3. one ncu capture of the prototype GEMM at the timing size (tensor pipe, L2 / DRAM traffic, stall reasons)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_i8_kernel -c 1 -o "$OUT/gram_i8_proto" \
  tools/gram_i8_proto 65536 5000 > "$OUT/ncu.txt" 2>&1
# 4. the whole GPU suite last: it is the longest step, and the round-end driver run repeats it anyway
timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.txt" 2>&1
echo "gpu suite exit $?" >> "$OUT/pytest_gpu.txt"
tail -5 "$OUT/pytest_new.txt" "$OUT/pytest_gpu.txt" "$OUT/gram_i8_proto.txt" "$OUT/k1_i8_proto.txt"
