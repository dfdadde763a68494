#!/bin/bash
# round 2, GPU call 19: how much of a K-heavy layer is weight traffic through L2 (half-B experiment, debug build)
set +e
mkdir -p gpurun_out
export SLQ_DEBUG_LIB=1
for L in "256 256 3 1 14" "1024 256 1 1 14" "512 512 3 1 7" "128 128 3 1 28" "2048 512 1 1 7"; do
  python tools/layer_time.py $L 256 2>&1 | tail -n 1
  SLQ_HALF_B=1 python tools/layer_time.py $L 256 2>&1 | tail -n 1 | sed 's/^/   half-B: /'
done
