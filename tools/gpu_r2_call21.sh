#!/bin/bash
# round 2, GPU call 21: residual unpack without the conversion unit (A/B against the I2F.U8 build), avgpool with pixel parts
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_conv_gpu.py tests/test_block_tail_gpu.py -x -q > gpurun_out/t_conv.log 2>&1; echo "conv+forward rc=$?"
tail -n 4 gpurun_out/t_conv.log
for cfg in "64 256 1 1 56 256 res" "128 512 1 1 28 256 res" "256 1024 1 1 14 256 res" "512 2048 1 1 7 256 res" "64 256 1 1 56 256 sres"; do
  python tools/layer_time.py $cfg 2>&1 | tail -n 1
  SLQ_LIB_VARIANT=resi2f python tools/layer_time.py $cfg 2>&1 | tail -n 1
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['conv_ms_per_step_serialised'], d['logits_rel_l2_vs_fp32'], d['top1_agreement_vs_fp32'])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include slq_step/ --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
grep -i "avgpool\|fc_\|stem" gpurun_out/launches.csv | cut -c1-200 | tail -n 8
