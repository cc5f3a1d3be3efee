#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_2gpubox.json 2>/dev/null
python -c "
import json
a=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/r2_bench_n1_2gpubox.json').read().strip().splitlines()[-1])
print('n2 ms', a['ms_per_step'], 'n1 ms', b['ms_per_step'], 'efficiency', b['ms_per_step']/a['ms_per_step'])"
