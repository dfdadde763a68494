#!/bin/bash
# round 2, GPU call 31: knock-out builds of the conv epilogue (timing only, wrong results): what each part costs
set +e
for cfg in "64 256 1 1 56 256 res" "128 512 1 1 28 256 res" "64 64 1 1 56 256"; do
  for v in "" koconst koi2f kores kopack koall; do
    SLQ_LIB_VARIANT=$v timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1
  done
done
