#!/bin/bash
# tools/gpu_profile.sh -- plain run, ncu launch list of ONE step (the NVTX range bench.py puts around its
# last timed step in --no-graph mode), then ncu --set full of the same step (reduced to a CSV on the box)
# and two full reports with source-level detail.
set +e
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"
[ $rc -ne 0 ] && exit 1
NV='--nvtx --nvtx-include slq_step/'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none $NV --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none $NV -o /tmp/prof_step $CMD \
    > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > /tmp/prof_step_raw.csv 2> gpurun_out/ncu_export.err
python tools/ncu_reduce.py /tmp/prof_step_raw.csv gpurun_out/prof_step_summary.csv; echo "reduce rc=$?"
python tools/ncu_summary.py gpurun_out/prof_step_summary.csv gpurun_out/ncu_full_step_summary.csv
# two conv launches with source-level detail: the first stage-1 expansion with residual (4th conv of the
# step) and a K-heavy 3x3 layer (29th)
for idx in 3 28; do
  timeout 600 ncu --set full --clock-control none --import-source on $NV -k regex:conv_umma -s $idx -c 1 \
      -o gpurun_out/prof_conv_op${idx} $CMD > gpurun_out/ncu_op${idx}.log 2>&1; echo "ncu op$idx rc=$?"
done
ls -la gpurun_out/ | head -40
