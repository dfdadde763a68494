#!/bin/bash
# round 2 evidence: full GPU test suite, bench lines, then (each only after the plain command exited 0) the ncu
# launch list of ONE step and the same step under `ncu --set full`, reduced to CSV on the box; one conv launch and
# one block-tail launch with source-level detail for the stall-site tables.
set +e
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -n 4 gpurun_out/t_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 50 --warmup 10 --layers gpurun_out/layers.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-agree --no-packed-b > gpurun_out/bench_nopacked.log 2> gpurun_out/bench_nopacked.err; echo "nopacked rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-agree --no-fuse-tail > gpurun_out/bench_nofuse.log 2> gpurun_out/bench_nofuse.err; echo "nofuse rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_ref", "bench_nopacked", "bench_nofuse"):  # (the side configurations are printed by hand)
    try:
        d = json.loads([l for l in open("gpurun_out/%s.log" % f) if l.startswith("{")][-1])
        print(f, d["value"], d.get("ms_per_step"), d.get("e2e", {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "failed", ex)
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log | cut -c1-200
timeout 300 python bench.py --arch resnet34 --batch 128 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r34.log 2> gpurun_out/bench_r34.err; echo "r34 rc=$?"
timeout 300 python bench.py --arch resnet18 --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r18.log 2> gpurun_out/bench_r18.err; echo "r18 rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-agree"
timeout 600 $CMD > gpurun_out/plain.log 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"
[ $rc -ne 0 ] && exit 1
NV='--nvtx --nvtx-include slq_step/'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none $NV --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none $NV -o /tmp/prof_step $CMD \
    > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > /tmp/prof_step_raw.csv 2> gpurun_out/ncu_export.err
python tools/ncu_reduce.py /tmp/prof_step_raw.csv gpurun_out/prof_step_summary.csv; echo "reduce rc=$?"
python tools/ncu_summary.py gpurun_out/prof_step_summary.csv gpurun_out/ncu_full_step_summary.csv
# source-level detail: the first stage-1 expansion with residual (3rd conv_umma launch after the tail), a K-heavy 3x3
# layer, and the stage-1 block tail
timeout 600 ncu --set full --clock-control none --import-source on $NV -k regex:conv_umma -s 4 -c 1 \
    -o gpurun_out/prof_conv_expand $CMD > gpurun_out/ncu_expand.log 2>&1; echo "ncu expand rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on $NV -k regex:conv_umma -s 22 -c 1 \
    -o gpurun_out/prof_conv_3x3 $CMD > gpurun_out/ncu_3x3.log 2>&1; echo "ncu 3x3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on $NV -k regex:block_tail -s 0 -c 1 \
    -o gpurun_out/prof_block_tail $CMD > gpurun_out/ncu_tail.log 2>&1; echo "ncu tail rc=$?"
ls -la gpurun_out/ | grep -i "ncu\|prof\|launch"
