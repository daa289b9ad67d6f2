#!/bin/bash
# Round measurement pass (1 GPU): tests, bench (both arms), ncu launch list + full capture of the two sweeps.
# Each ncu command runs only after its plain command exited 0.  Usage: tools/run_measure.sh TAG
T=${1:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain_$T.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_density_persist|k_density_tiled|k_force_stream" -s 6 -c 2 \
    -o gpurun_out/prof_$T -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_$T.log 2>&1
ncu -i gpurun_out/prof_$T.ncu-rep --page raw --csv > gpurun_out/raw_$T.csv
python -c "import json; d=json.loads(open('gpurun_out/bench_$T.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['kernel'][:20], d['roofline']['frac'], d['cpu_baseline']['value'])"
