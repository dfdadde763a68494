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
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"conv_umma|stem_|avgpool|fc_kernel" -s 284 -c 57 --csv --log-file gpurun_out/launches.csv $CMD \
    > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
tail -n 3 gpurun_out/t_*.log gpurun_out/smoke.log gpurun_out/bench.log
grep DIAG gpurun_out/diag_conv.log | cut -c1-400
