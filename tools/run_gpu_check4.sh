python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "--- config 1 with graph replay"; python tools/config1_default_scene.py 2>&1 | head -3
echo "--- config 1 without"; SPHB200_NO_GRAPH=1 python tools/config1_default_scene.py 2>&1 | head -2
for w in dambreak_1m dambreak_16m; do for g in 0 1; do
SPHB200_NO_GRAPH=$g python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_g$g.json').read()); print('$w nograph=$g', d['ms_per_step'], d['gpu_launches'])"
done; done
