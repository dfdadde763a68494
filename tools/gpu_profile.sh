#!/bin/bash
# tools/gpu_profile.sh -- plain run, ncu launch list of one step, then ncu --set full captures of the
# stem and of every conv launch of one step (reduced to a CSV on the box; two full reports kept).
set +e
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"
[ $rc -ne 0 ] && exit 1
# launches before the timed steps: calibrate (1+1 stem, 52*2 conv) + 1 forward (56) + 3 warm-up (168) = 330
KREG='regex:conv_umma|stem_|avgpool|fc_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 166 -c 56 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none -k "$KREG" -s 166 -c 56 -o /tmp/prof_step $CMD \
    > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > /tmp/prof_step_raw.csv 2> gpurun_out/ncu_export.err
python tools/ncu_reduce.py /tmp/prof_step_raw.csv gpurun_out/prof_step_summary.csv; echo "reduce rc=$?"
# two launches with source-level detail: a stage-1 1x1+residual layer and a K-heavy 3x3 layer
for idx in 3 28; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s $((164+idx)) -c 1 \
      -o gpurun_out/prof_conv_op${idx} $CMD > gpurun_out/ncu_op${idx}.log 2>&1; echo "ncu op$idx rc=$?"
done
ls -la gpurun_out/ | head -40
