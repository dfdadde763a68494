#!/bin/bash
# round 2, GPU call 32: residual conversion split between the XU pipe and PRMT (A/B)
set +e
for cfg in "64 256 1 1 56 256 res" "128 512 1 1 28 256 res" "256 1024 1 1 14 256 res" "512 2048 1 1 7 256 res"; do
  for v in "" ressplit "" ressplit; do
    SLQ_LIB_VARIANT=$v timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1
  done
done
