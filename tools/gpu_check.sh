#!/bin/bash
# tools/gpu_check.sh -- GPU parity suite, short bench, then the per-launch ncu time list of one step.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ] && [ "$1" != "nolist" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k regex:"conv_umma|stem_|avgpool|fc_kernel" -s 166 -c 56 --csv --log-file gpurun_out/launches.csv $CMD \
      > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
fi
cat gpurun_out/bench.log
