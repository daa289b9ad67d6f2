#!/bin/bash
# 8-GPU box: weak (16.7M per GPU) and strong (16.7M total) scaling bench lines
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench8_weak_r01d.json 2> gpurun_out/bench8_weak_r01d.err
python -c "import json; d=json.loads(open('gpurun_out/bench8_weak_r01d.json').read()); print('weak8', d['ms_per_step'], '%.3e' % d['value'], '%.3e' % d['e2e']['value'], d['config']['phase_ms_rank0'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --scaling strong > gpurun_out/bench8_strong_r01d.json 2> gpurun_out/bench8_strong_r01d.err
python -c "import json; d=json.loads(open('gpurun_out/bench8_strong_r01d.json').read()); print('strong8', d['ms_per_step'], '%.3e' % d['value'], d['config']['phase_ms_rank0'])"
