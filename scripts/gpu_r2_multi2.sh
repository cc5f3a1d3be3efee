#!/bin/bash
# phase breakdown of sharded resample steps + bench at G GPUs
G=${1:-8}
mkdir -p gpurun_out
for thr in 0.1 1.0; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29521 \
  scripts/phase_timing_multi.py 20 200 $thr 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" ; done | tee gpurun_out/r2_phases_$G.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r2e_bench_n$G.json 2> gpurun_out/r2e_bench_n$G.err
python -c "import json;d=json.load(open('gpurun_out/r2e_bench_n$G.json'));print('bench n=$G ms',d['ms_per_step'],'value',d['value'])"
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_n1_same_box.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r2e_bench_n1_same_box.json'));print('bench n=1 ms',d['ms_per_step'],'value',d['value'])"
