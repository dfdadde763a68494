#!/bin/bash
# round 2, 8-GPU call (charged 8x): the sensitivity sweep at 8 and 4 GPUs and the bench at N = 8, 4.
# The 1- and 2-GPU legs run in cheaper calls (tools/gpu_r2_scale2.sh); identical values digest expected at every N.
set +e
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
  timeout 240 $TR --nproc-per-node $N --master-port $((29500 + N)) tools/sweep_run.py --arch resnet50 > gpurun_out/sweep_$N.json 2> gpurun_out/sweep_$N.err; echo "sweep $N rc=$?"
  tail -n 1 gpurun_out/sweep_$N.json | cut -c1-330
done
for N in 8 4; do
  timeout 300 $TR --nproc-per-node $N --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-agree > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench $N rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n$N.log') if l.startswith('{')][-1])
    print($N, {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'probe', d['e2e'].get('h2d_probe_gbs_per_gpu'), d['e2e'].get('numa'), d['clocks'])
except Exception as ex:
    print($N, 'failed', ex)
PY
done
