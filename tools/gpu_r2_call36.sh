#!/bin/bash
# round 2, GPU call 36: one staging tile per team against two (A/B) on the layers that regressed against round 1
set +e
for cfg in "64 256 1 1 56 256 res" "256 1024 1 1 14 256 res" "512 2048 1 1 7 256 res" "256 128 1 1 56 256" "64 64 1 1 56 256" "512 128 1 1 28 256"; do
  LT_NO_RS=1 timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1
  LT_NO_RS=1 SLQ_LIB_VARIANT=sstg timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1
done
