#!/bin/bash
# round 2, GPU call 29: sweep on the caller's net (no work model) vs the general path; timing
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep_gpu.py tests/test_reference_mains_gpu.py tests/test_eval_store_gpu.py -x -q > gpurun_out/t_sweep.log 2>&1; echo "sweep tests rc=$?"
tail -n 4 gpurun_out/t_sweep.log | cut -c1-300
timeout 240 python tools/sweep_run.py --arch resnet50 > gpurun_out/sweep_1.json 2> gpurun_out/sweep_1.err; echo "sweep rc=$?"
tail -n 1 gpurun_out/sweep_1.json | cut -c1-420
timeout 240 python tools/sweep_run.py --arch resnet50 --force-work-model > gpurun_out/sweep_1w.json 2> gpurun_out/sweep_1w.err; echo "sweep (work model) rc=$?"
tail -n 1 gpurun_out/sweep_1w.json | cut -c1-420
timeout 240 python tools/sweep_run.py --arch resnet50 --trace-evals > /dev/null 2> gpurun_out/sweep_1t.err; grep "^eval" gpurun_out/sweep_1t.err | head -n 6
