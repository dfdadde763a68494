#!/bin/bash
# round 2, GPU call 35: output channel sums (dp4a per four outputs) only where a consumer gathers them
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py -x -q > gpurun_out/t_conv.log 2>&1; echo "conv+forward rc=$?"
tail -n 3 gpurun_out/t_conv.log | cut -c1-300
for cfg in "64 256 1 1 56 256 res" "128 512 1 1 28 256 res" "64 64 1 1 56 256" "256 64 1 1 56 256"; do
  timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1
  LT_NO_RS=1 timeout 120 python tools/layer_time.py $cfg 2>&1 | tail -n 1 | sed 's/^prod/no-rs/'
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['conv_ms_per_step_serialised'], d['logits_rel_l2_vs_fp32'], d['top1_agreement_vs_fp32'])
PY
sed -n 1,10p gpurun_out/layers.txt
