#!/bin/bash
# neighbour-density sweep at 16.7M (config 5) + the 1M dam-break (config 2), 1 GPU
for nu in 30 60 120; do
python bench.py --nu $nu --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nu$nu.json 2> gpurun_out/bench_nu$nu.err
python -c "import json; d=json.loads(open('gpurun_out/bench_nu$nu.json').read()); print('nu=$nu', round(d['ms_per_step'],3), '%.3e' % d['value'], '%.3e' % d['config']['neighbor_pairs_per_sec'], round(d['config']['mean_neighbors'],1), d['config']['phase_ms_rank0'])"
done
python bench.py --workload dambreak_1m --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
python -c "import json; d=json.loads(open('gpurun_out/bench_1m.json').read()); print('1m', round(d['ms_per_step'],4), '%.3e' % d['value'], '%.3e' % d['config']['neighbor_pairs_per_sec'], '%.3e' % d['e2e']['value'])"
