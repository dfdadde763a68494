#!/bin/bash
# round 2, GPU call 30: 256-channel (wide) against 128-channel tiles on the small-M K-heavy layers (debug build knob)
set +e
export SLQ_DEBUG_LIB=1
for cfg in "2048 512 1 1 7 256" "512 512 3 1 7 256" "512 512 3 2 14 256" "1024 512 1 1 14 256" "1024 256 1 1 14 256" "256 256 3 1 14 256" "256 256 3 2 28 256"; do
  timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1
  SLQ_NO_WIDE=1 timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1 | sed 's/^prod/narrow/'
done
