#!/bin/bash
# A/B timing of library variants (tools/build_variant.py): tools/ab_variants.sh name1 name2 ...
for v in "$@"; do
  lib=smoothed_particle_hydrodynamics_b200/variants/libsphb200_$v.so
  SPHB200_LIB=$PWD/$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/ab_$v.json').read()); print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['config']['phase_ms_rank0'].items()})"
done
