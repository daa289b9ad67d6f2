#!/bin/bash
# neighbour-density sweep at 16.7M particles on one GPU (config 5): tools/run_sweep.sh
for nu in 40 60 120; do
  python bench.py --steps 10 --warmup 3 --nu $nu --no-extras --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_nu$nu.json
  python -c "import json; d=json.loads(open('gpurun_out/sweep_nu$nu.json').read()); print('nu $nu', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['config']['phase_ms_rank0'].items()})"
done
