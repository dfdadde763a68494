#!/bin/bash
# tools/gpu_quick.sh -- conv parity tests, then a short bench with the per-layer table (no CPU baseline).
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q --timeout=300 -p no:cacheprovider 2>&1 | grep -v "mbarrier timeout" | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --layers gpurun_out/layers.txt --no-cpu-baseline "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?"
grep -v "mbarrier timeout" gpurun_out/bench.log | cut -c1-330
grep -c "mbarrier timeout" gpurun_out/bench.log
tail -3 gpurun_out/bench.err
