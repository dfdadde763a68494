#!/bin/bash
# round 2, GPU call 2: conv v2 parity (conv + forward tests first), whole suite, bench line with per-layer table
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py -x -q > gpurun_out/t_conv.log 2>&1; echo "conv+forward rc=$?"
tail -n 12 gpurun_out/t_conv.log
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_conv_gpu.py --deselect tests/test_forward_gpu.py > gpurun_out/t_gpu.log 2>&1; echo "rest rc=$?"
tail -n 25 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -v "mbarrier timeout" gpurun_out/bench.err | tail -c 1500
cat gpurun_out/bench.log
cat gpurun_out/layers.txt
