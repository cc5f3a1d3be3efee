#!/bin/bash
# sharded-filter validation + weak-scaling timing on however many GPUs the box has
G=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
echo "== multi_gpu_worker on $G GPUs"
timeout 600 python -m torch.distributed.run --standalone --nnodes=1 --nproc-per-node $G tests/multi_gpu_worker.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/multi_gpu_worker_$G.log
echo "== timing"
timeout 600 python -m torch.distributed.run --standalone --nnodes=1 --nproc-per-node $G scripts/multi_gpu_timing.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/multi_gpu_timing_$G.log
echo "== bench N=$G"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$G.json
