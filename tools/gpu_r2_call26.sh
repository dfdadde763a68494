#!/bin/bash
# round 2, GPU call 26: activations from one arena, no-op net.to() skipped, resident in-memory loader: tests + sweep timing
set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -n 4 gpurun_out/t_gpu.log | cut -c1-300
timeout 240 python tools/sweep_run.py --arch resnet50 --trace-evals > gpurun_out/sweep_1t.json 2> gpurun_out/sweep_1t.err; echo "sweep rc=$?"
tail -n 1 gpurun_out/sweep_1t.json | cut -c1-260
grep "^eval" gpurun_out/sweep_1t.err | head -n 8
timeout 240 python tools/sweep_run.py --arch resnet50 > gpurun_out/sweep_1.json 2> gpurun_out/sweep_1.err; echo "sweep rc=$?"
tail -n 1 gpurun_out/sweep_1.json | cut -c1-260
