#!/bin/bash
# round 2, GPU call 4: tensor-core tail, rowsum reduction moved to the end of the tile, stem v2 timing breakdown
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_conv_gpu.py -x -q -s > gpurun_out/t_conv.log 2>&1; echo "conv+forward rc=$?"
grep -E "tail N=|passed|failed|Error" gpurun_out/t_conv.log | tail -n 12
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -v "mbarrier timeout" gpurun_out/bench.err | tail -c 1500
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['conv_ms_per_step_serialised'], d['logits_rel_l2_vs_fp32'], d['top1_agreement_vs_fp32'])
PY
cat gpurun_out/layers.txt
STEM_DBG_LIST="0,1,2,3,4,8,12,15" python tools/time_stem.py > gpurun_out/time_stem.log 2>&1; cat gpurun_out/time_stem.log | tail -n 10
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "slq_step/" --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
  python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5 and r[0].isdigit()]
tot=0
for r in rows:
    name=r[4][:60]; val=float(r[-1]); unit=r[-2]
    tot+=val
    print("%-60s %10.1f %s" % (name, val, unit))
print("total", tot)
PY
fi
