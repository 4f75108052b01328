#!/bin/bash
# Round 2, call 1: run the int8 digit-slice prototypes that round 1 left unrun (prebuilt binaries travel with the snapshot).
set -u
OUT=gpurun_out/r2_call1
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1
for args in "4096 256" "65536 640" "131072 5000"; do
  echo "== gram $args" >> "$OUT/gram_i8_proto.txt"
  timeout 120 tools/gram_i8_proto $args >> "$OUT/gram_i8_proto.txt" 2>&1
  echo "exit $?" >> "$OUT/gram_i8_proto.txt"
done
echo "== gram 65536 640, 7-bit digits" >> "$OUT/gram_i8_proto.txt"
timeout 120 tools/gram_i8_proto7 65536 640 >> "$OUT/gram_i8_proto.txt" 2>&1
echo "exit $?" >> "$OUT/gram_i8_proto.txt"
for args in "4096 256" "131072 5000"; do
  echo "== gram $args, CLUSTER=2" >> "$OUT/gram_i8_proto.txt"
  timeout 120 tools/gram_i8_proto_c2 $args >> "$OUT/gram_i8_proto.txt" 2>&1
  echo "exit $?" >> "$OUT/gram_i8_proto.txt"
done
for args in "1024 256" "8192 640" "131072 2560"; do
  echo "== trsm $args" >> "$OUT/gram_i8_proto.txt"
  timeout 180 tools/gram_i8_proto trsm $args >> "$OUT/gram_i8_proto.txt" 2>&1
  echo "exit $?" >> "$OUT/gram_i8_proto.txt"
done
for exe in k1_i8_proto_v3 k1_i8_proto_v2; do
  echo "== $exe 131072 5120" >> "$OUT/k1_i8_proto.txt"
  timeout 120 tools/$exe 131072 5120 >> "$OUT/k1_i8_proto.txt" 2>&1
  echo "exit $?" >> "$OUT/k1_i8_proto.txt"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_i8_kernel -c 1 -o "$OUT/k1_i8_proto_v3" \
  tools/k1_i8_proto_v3 131072 5120 > "$OUT/ncu_k1.txt" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram_i8_kernel -c 1 -o "$OUT/gram_i8_proto" \
  tools/gram_i8_proto 65536 5000 > "$OUT/ncu.txt" 2>&1
tail -30 "$OUT/gram_i8_proto.txt" "$OUT/k1_i8_proto.txt"
