#!/bin/bash
# round 2: the sensitivity sweep at 8 and 4 GPUs with the final code (1 and 2 GPUs: tools/gpu_r2_scale2.sh)
set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
  timeout 240 $TR --nproc-per-node $N --master-port $((29500 + N)) tools/sweep_run.py --arch resnet50 > gpurun_out/sweep_$N.json 2> gpurun_out/sweep_$N.err; echo "sweep $N rc=$?"
  tail -n 1 gpurun_out/sweep_$N.json | cut -c1-330
done
