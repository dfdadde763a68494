#!/bin/bash
# tools/gpu_r2_call1.sh -- round 2, first GPU call: full GPU suite, bench line, quantizer ncu capture, 1-GPU sweep.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
lscpu | head -20 > gpurun_out/lscpu.txt; numactl -H >> gpurun_out/lscpu.txt 2>&1
nvidia-smi topo -m >> gpurun_out/lscpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.err
cat gpurun_out/bench.log
timeout 300 python tools/sweep_run.py --arch resnet50 --batches 2 --batch 64 > gpurun_out/sweep_w1.log 2> gpurun_out/sweep_w1.err; echo "sweep rc=$?"
tail -n 2 gpurun_out/sweep_w1.log; tail -n 5 gpurun_out/sweep_w1.err
CMD="python tools/quant_run.py resnet50 3"
timeout 300 $CMD > gpurun_out/quant_plain.log 2>&1
rc=$?; echo "quant plain rc=$rc"; cat gpurun_out/quant_plain.log | tail -4
if [ $rc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:quantize_jobs -s 1 -c 2 \
      -o gpurun_out/r2_prof_quantizer $CMD > gpurun_out/ncu_quant.log 2>&1; echo "ncu quant rc=$?"
fi
