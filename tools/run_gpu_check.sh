python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p6.json 2> gpurun_out/bench_p6.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_p6.json').read()); print(d['ms_per_step'], d['config']['phase_ms_rank0'])"
ncu --set full --clock-control none --import-source on -k regex:k_density_tiled -s 2 -c 1 -o gpurun_out/prof_r1l -f python bench.py --workload dambreak_1m --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu -i gpurun_out/prof_r1l.ncu-rep --page raw --csv > gpurun_out/raw_l.csv
ncu -i gpurun_out/prof_r1l.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/src_density_l.csv 2>/dev/null
ls -la gpurun_out/prof_r1l.ncu-rep
