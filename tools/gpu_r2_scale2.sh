#!/bin/bash
# round 2, 2-GPU call: the sweep at 1 and 2 GPUs (same digest as tools/gpu_r2_scale8.sh expected) and the bench at N = 2
set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 python tools/sweep_run.py --arch resnet50 > gpurun_out/sweep_1.json 2> gpurun_out/sweep_1.err; echo "sweep 1 rc=$?"
timeout 240 $TR --nproc-per-node 2 --master-port 29502 tools/sweep_run.py --arch resnet50 > gpurun_out/sweep_2.json 2> gpurun_out/sweep_2.err; echo "sweep 2 rc=$?"
for N in 1 2; do tail -n 1 gpurun_out/sweep_$N.json | cut -c1-330; done
timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-agree > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench 2 rc=$?"
grep '^{' gpurun_out/bench_n2.log | tail -1 | cut -c1-400
