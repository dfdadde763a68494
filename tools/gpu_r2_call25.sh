#!/bin/bash
# round 2, GPU call 25: where the sweep's fixed cost goes (cProfile of a short sweep)
set +e
mkdir -p gpurun_out
timeout 300 python -m cProfile -o gpurun_out/sweep.prof tools/sweep_run.py --arch resnet50 --layers 9 > gpurun_out/sweep_prof.json 2> gpurun_out/sweep_prof.err; echo "rc=$?"
python - <<'PY'
import pstats
p = pstats.Stats('gpurun_out/sweep.prof')
p.sort_stats('cumulative').print_stats(70)
PY
