#!/bin/bash
# round 2: ncu launch list of ONE step of the final commit (after the plain command has exited 0)
set +e
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 300 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err; rc=$?; echo "plain rc=$rc"; [ $rc -ne 0 ] && exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include slq_step/ --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/launches.csv
