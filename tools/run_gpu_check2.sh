python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for r in 1 0; do
SPHB200_RADIX_SORT=$r python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sort$r.json 2> gpurun_out/bench_sort$r.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_sort$r.json').read()); print('radix=$r', d['ms_per_step'], d['config']['phase_ms_rank0'])"
done
