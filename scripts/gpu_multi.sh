#!/bin/bash
# sharded-filter validation + weak-scaling evidence on however many GPUs the box has: bench.py at N = G, G/2, ... , 1
G=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
echo "== multi_gpu_worker on $G GPUs"
timeout 120 python -m torch.distributed.run --standalone --nnodes=1 --nproc-per-node $G tests/multi_gpu_worker.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/multi_gpu_worker_$G.log
echo "== timing (configs 2, 4, 5 shapes per GPU)"
timeout 90 python -m torch.distributed.run --standalone --nnodes=1 --nproc-per-node $G scripts/multi_gpu_timing.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee gpurun_out/multi_gpu_timing_$G.log
n=$G
while [ $n -ge 2 ]; do
echo "== bench N=$n"
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$n.json
n=$((n/2))
done
echo "== bench N=1"; timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n1_samebox.json
