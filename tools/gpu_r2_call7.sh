#!/bin/bash
set +e
mkdir -p gpurun_out
python tools/repro_small.py resnet50 64 > gpurun_out/repro_plain.log 2>&1; echo "plain rc=$?"; tail -n 12 gpurun_out/repro_plain.log
python tools/repro_small.py resnet18 224 > gpurun_out/repro_plain18.log 2>&1; echo "plain18 rc=$?"; tail -n 8 gpurun_out/repro_plain18.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 30 python tools/repro_small.py resnet50 64 > gpurun_out/repro_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "Invalid|misaligned|Misaligned|at 0x|by thread|Address|ERROR SUMMARY|kernel" gpurun_out/repro_memcheck.log | head -40
