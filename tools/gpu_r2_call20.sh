#!/bin/bash
# round 2, GPU call 20: timeline + wait statistics of a K-heavy layer (3x3 256->256 @14, batch 256)
set +e
mkdir -p gpurun_out
timeout 300 python tools/wait_stats.py 256 256 3 1 14 256 > gpurun_out/wait_stats_256.log 2>&1; tail -n 14 gpurun_out/wait_stats_256.log
timeout 300 python tools/wait_stats.py 1024 256 1 1 14 256 > gpurun_out/wait_stats_1024.log 2>&1; tail -n 14 gpurun_out/wait_stats_1024.log
TRACE_FROM=200 TRACE_TO=330 timeout 300 python tools/trace_conv.py 256 256 3 1 14 256 > gpurun_out/trace_conv_256.log 2>&1; tail -n 140 gpurun_out/trace_conv_256.log
