#!/bin/bash
# round 2, GPU call 5: stem v2 with one 3-D TMA per pair + register vertical max; FFMA2 A/B; ncu source profile of conv op 3
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_conv_gpu.py -x -q > gpurun_out/t_conv.log 2>&1; echo "conv+forward rc=$?"
tail -n 6 gpurun_out/t_conv.log
for cfg in "64 256 1 1 56 256 res" "64 256 1 1 56 256 w16" "128 512 1 1 28 256 res" "64 64 3 1 56 256" "64 64 1 1 56 256" "256 1024 1 1 14 256 res"; do
  python tools/layer_time.py $cfg 2>&1 | tail -n 1
  SLQ_LIB_VARIANT=noffma2 python tools/layer_time.py $cfg 2>&1 | tail -n 1
done
STEM_DBG_LIST="0,4,8,12,15" python tools/time_stem.py 2>&1 | tail -n 6
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -v "mbarrier timeout" gpurun_out/bench.err | tail -c 1500
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['conv_ms_per_step_serialised'], d['logits_rel_l2_vs_fp32'], d['top1_agreement_vs_fp32'])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "slq_step/" -k regex:conv_umma -s 3 -c 1 \
      -o gpurun_out/r2_prof_conv_op3 $CMD > gpurun_out/ncu_op3.log 2>&1; echo "ncu op3 rc=$?"
fi
