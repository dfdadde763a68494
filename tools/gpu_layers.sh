#!/bin/bash
# tools/gpu_layers.sh tag [ENV=VALUE ...] -- per-layer table of one bench run under the given environment
tag=$1; shift
env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --layers gpurun_out/layers_$tag.txt --no-cpu-baseline --no-agree > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
echo "$tag rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_$tag.log | head -1)"
