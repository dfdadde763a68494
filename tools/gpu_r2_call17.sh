#!/bin/bash
# round 2, GPU call 17-18: block tail wait statistics per role and crew segments
set +e
mkdir -p gpurun_out
BT_DBG_LIST="0,47" timeout 600 python tools/time_tail.py > gpurun_out/time_tail.log 2>&1; echo "time_tail rc=$?"
cat gpurun_out/time_tail.log
