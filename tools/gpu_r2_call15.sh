#!/bin/bash
# round 2, GPU call 15: where the block tail's time goes (debug-build knobs)
set +e
mkdir -p gpurun_out
timeout 600 python tools/time_tail.py > gpurun_out/time_tail.log 2>&1; echo "time_tail rc=$?"
cat gpurun_out/time_tail.log
