#!/bin/bash
# round 2: the sensitivity sweep at 8 GPUs (the 4-GPU run of tools/gpu_r2_scale8.sh and the 1-GPU run gave the same digest)
set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 2; do
  timeout 240 $TR --nproc-per-node $N --master-port $((29500 + N)) tools/sweep_run.py --arch resnet50 > gpurun_out/sweep_$N.json 2> gpurun_out/sweep_$N.err; echo "sweep $N rc=$?"
  tail -n 1 gpurun_out/sweep_$N.json | cut -c1-330
done
