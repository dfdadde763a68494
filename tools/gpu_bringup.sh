#!/bin/bash
# tools/gpu_bringup.sh -- one gpurun call: staged GPU checks, each isolated in its own process.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
PT="python -m pytest -q --timeout=600 -p no:cacheprovider"
timeout 900 $PT tests/test_quantizer_gpu.py > gpurun_out/t_quant.log 2>&1; echo "quant rc=$?"
timeout 600 $PT tests/test_conv_gpu.py -k "gemm_ready or simt_checker" > gpurun_out/t_simt.log 2>&1; echo "simt rc=$?"
timeout 900 python tools/diag_conv.py > gpurun_out/diag_conv.log 2>&1; echo "diag rc=$?"
timeout 1200 $PT tests/test_conv_gpu.py -k "umma" > gpurun_out/t_umma.log 2>&1; echo "umma rc=$?"
timeout 1200 $PT tests/test_forward_gpu.py > gpurun_out/t_forward.log 2>&1; echo "forward rc=$?"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --simt --no-cpu-baseline > gpurun_out/bench_simt.log 2> gpurun_out/bench_simt.err; echo "bench_simt rc=$?"
tail -n 3 gpurun_out/t_*.log gpurun_out/smoke.log gpurun_out/bench.log
grep DIAG gpurun_out/diag_conv.log | cut -c1-400
