#!/bin/bash
# tools/gpu_profile.sh -- plain run, then ncu launch list + full-set capture of the conv kernels.
# Big .ncu-rep files are reduced to CSV on the box (gpurun_out/ is capped at 64 MiB).
set +e
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/diag_*.npz
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
      -o /tmp/prof_convs $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu -i /tmp/prof_convs.ncu-rep --page raw --csv > /tmp/prof_convs_raw.csv 2> gpurun_out/ncu_export.err
  python tools/ncu_reduce.py /tmp/prof_convs_raw.csv gpurun_out/prof_convs_summary.csv; echo "reduce rc=$?"
  for idx in 1 28; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s $((260+idx)) -c 1 \
        -o gpurun_out/prof_conv_op${idx} $CMD > gpurun_out/ncu_op${idx}.log 2>&1; echo "ncu op$idx rc=$?"
  done
  ls -la gpurun_out/
fi
tail -n 14 gpurun_out/diag_div.log
