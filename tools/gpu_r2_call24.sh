#!/bin/bash
# round 2, GPU call 24: stores straight from registers (no staging tile / team barrier / TMA store) on the resident-weight
# layers, A/B; the sweep with its evaluations timed one by one
set +e
mkdir -p gpurun_out
for cfg in "64 256 1 1 56 256 res" "128 512 1 1 28 256 res" "256 1024 1 1 14 256 res" "512 2048 1 1 7 256 res" "64 64 1 1 56 256" "64 64 3 1 56 256" "256 64 1 1 56 256" "128 128 3 1 28 256" "512 128 1 1 28 256"; do
  python tools/layer_time.py $cfg 2>&1 | tail -n 1
  SLQ_LIB_VARIANT=direct timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1
done
SLQ_LIB_VARIANT=direct timeout 600 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -n 3
timeout 240 python tools/sweep_run.py --arch resnet50 --trace-evals > gpurun_out/sweep_1t.json 2> gpurun_out/sweep_1t.err; echo "sweep rc=$?"
tail -n 1 gpurun_out/sweep_1t.json | cut -c1-260
grep "^eval" gpurun_out/sweep_1t.err | head -n 14
grep "^eval" gpurun_out/sweep_1t.err | tail -n 3
