#!/bin/bash
# round 2, GPU call 9: ones row back for narrow tiles, rowsum planes only for 256-channel tiles
set +e
mkdir -p gpurun_out
python tools/repro_small.py resnet50 64 > gpurun_out/repro_plain.log 2>&1; echo "repro rc=$?"; tail -n 6 gpurun_out/repro_plain.log
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_forward_gpu.py -x -q > gpurun_out/t_conv.log 2>&1; echo "conv+forward rc=$?"
tail -n 6 gpurun_out/t_conv.log
STEM_DBG_LIST="0,4,8,12,15" python tools/time_stem.py > gpurun_out/time_stem.log 2>&1; tail -n 6 gpurun_out/time_stem.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -v "mbarrier timeout" gpurun_out/bench.err | tail -c 1500
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['conv_ms_per_step_serialised'], d['logits_rel_l2_vs_fp32'], d['top1_agreement_vs_fp32'])
PY
cat gpurun_out/layers.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-agree --no-packed-b --layers gpurun_out/layers_nopacked.txt > gpurun_out/bench_nopacked.log 2> gpurun_out/bench_nopacked.err; echo "bench nopacked rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_nopacked.log').read().strip().splitlines()[-1])
print("no packed:", {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['conv_ms_per_step_serialised'])
PY
head -16 gpurun_out/layers_nopacked.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_conv_gpu.py --deselect tests/test_forward_gpu.py > gpurun_out/t_gpu.log 2>&1; echo "rest rc=$?"
tail -n 5 gpurun_out/t_gpu.log
