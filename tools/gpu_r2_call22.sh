#!/bin/bash
# round 2, GPU call 22: TMEM read ceiling, epilogue segment timing, setmaxnreg A/B, 1-GPU sweep digest
set +e
mkdir -p gpurun_out
timeout 120 tools/bin/ldtm_bench > gpurun_out/ldtm_bench.log 2>&1; echo "ldtm rc=$?"; cat gpurun_out/ldtm_bench.log
for cfg in "64 256 1 1 56 256 res" "128 512 1 1 28 256 res" "256 1024 1 1 14 256 res" "64 64 1 1 56 256" "64 64 3 1 56 256" "256 64 1 1 56 256"; do
  timeout 120 python tools/wait_stats.py $cfg 2>&1 | tail -n 18
done > gpurun_out/wait_stats_epi.log
cat gpurun_out/wait_stats_epi.log
for cfg in "64 256 1 1 56 256 res" "256 1024 1 1 14 256 res" "64 64 3 1 56 256" "256 256 3 1 14 256" "256 64 1 1 56 256"; do
  python tools/layer_time.py $cfg 2>&1 | tail -n 1
  SLQ_LIB_VARIANT=smr timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1
done
timeout 240 python tools/sweep_run.py --arch resnet50 > gpurun_out/sweep_1.json 2> gpurun_out/sweep_1.err; echo "sweep 1 rc=$?"
tail -n 1 gpurun_out/sweep_1.json | cut -c1-330
