#!/bin/bash
# round 2, GPU call 3: stem v2 (TMEM A operand) parity, conv epilogue fix, wait statistics of the 256-channel tiles
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -k "stem or u8_and_fp16 or first_block" > gpurun_out/t_stem.log 2>&1; echo "stem tests rc=$?"
tail -n 15 gpurun_out/t_stem.log
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/t_conv.log 2>&1; echo "conv+forward rc=$?"
tail -n 8 gpurun_out/t_conv.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -v "mbarrier timeout" gpurun_out/bench.err | tail -c 1500
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['conv_ms_per_step_serialised'])
PY
cat gpurun_out/layers.txt
python tools/time_stem.py > gpurun_out/time_stem.log 2>&1; SLQ_STEM_OLD=1 python tools/time_stem.py >> gpurun_out/time_stem.log 2>&1; tail -n 6 gpurun_out/time_stem.log
for cfg in "256 256 3 1 14" "1024 256 1 1 14" "512 512 3 1 7" "64 256 1 1 56"; do
  python tools/wait_stats.py $cfg 256 >> gpurun_out/wait_stats.log 2>&1
  SLQ_NO_WIDE=1 python tools/wait_stats.py $cfg 256 >> gpurun_out/wait_stats_nowide.log 2>&1
done
cat gpurun_out/wait_stats.log; echo ---- no wide; cat gpurun_out/wait_stats_nowide.log
