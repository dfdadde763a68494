#!/bin/bash
# round 2, GPU call 14: stem builders as two alternating groups; block tail back to the LDS-constants crew
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_block_tail_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/t_fwd.log 2>&1; echo "forward+tail+baseline rc=$?"
tail -n 5 gpurun_out/t_fwd.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -v "mbarrier timeout" gpurun_out/bench.err | tail -c 1500
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['conv_ms_per_step_serialised'], d['logits_rel_l2_vs_fp32'], d['top1_agreement_vs_fp32'], d['gpu_launches'])
PY
grep "tail\|  1 1  56  802816 0 1\|idx\| 128    0  512 1" gpurun_out/layers.txt
STEM_DBG_LIST="0,4" python tools/time_stem.py > gpurun_out/time_stem.log 2>&1; tail -n 3 gpurun_out/time_stem.log
timeout 600 python tools/trace_stem.py > gpurun_out/trace_stem.log 2>&1; echo "trace rc=$?"
tail -n 36 gpurun_out/trace_stem.log
