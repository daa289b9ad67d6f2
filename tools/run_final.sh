#!/bin/bash
# what the driver runs at round end, on one GPU: build check is done on the CPU box; here tests, smoke, both bench arms
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read())
print(d['ms_per_step'], '%.4g'%d['value'], '%.4g'%d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'][:16], round(d['roofline']['frac'],4), '%.4g'%d['cpu_baseline']['value'], d['clocks'])
r=json.loads(open('gpurun_out/bench_ref_final.json').read()); print('ref', '%.4g'%r['value'], r['cpu_baseline']['cores'])"
