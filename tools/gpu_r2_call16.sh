#!/bin/bash
# round 2, GPU call 16: one mbarrier arrival per epilogue warp (conv + block tail); stem with the BN scale folded into the weights
set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py tests/test_block_tail_gpu.py -x -q > gpurun_out/t_conv.log 2>&1; echo "conv+forward+tail rc=$?"
tail -n 8 gpurun_out/t_conv.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -v "mbarrier timeout" gpurun_out/bench.err | tail -c 1500
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['conv_ms_per_step_serialised'], d['logits_rel_l2_vs_fp32'], d['top1_agreement_vs_fp32'], d['gpu_launches'])
PY
cat gpurun_out/layers.txt
STEM_DBG_LIST="0,4" python tools/time_stem.py > gpurun_out/time_stem.log 2>&1; tail -n 3 gpurun_out/time_stem.log
BT_DBG_LIST="0,47" python tools/time_tail.py > gpurun_out/time_tail.log 2>&1; cat gpurun_out/time_tail.log
