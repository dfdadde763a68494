#!/bin/bash
# tools/gpu_profile.sh -- plain run, then ncu launch list + full-set capture of the conv kernels.
set +e
mkdir -p gpurun_out
timeout 300 python tools/diag_div.py > gpurun_out/diag_div.log 2>&1; echo "diag_div rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  # calibrate (52 conv + 2 stem) + 1 forward (56) + 3 warm-up steps (168) = 278 matching launches
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k regex:"conv_umma|stem_|avgpool|fc_kernel" -s 278 -c 56 --csv --log-file gpurun_out/launches.csv $CMD \
      > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
  timeout 1500 ncu --set full --clock-control none -k regex:conv_umma -s 260 -c 52 \
      -o gpurun_out/prof_convs $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ls -la gpurun_out/*.ncu-rep
fi
tail -n 12 gpurun_out/diag_div.log
