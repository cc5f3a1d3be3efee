#!/bin/bash
# round 2: sharded-filter parity + timing on G GPUs of one box:  gpurun --gpus G -- 'bash scripts/gpu_r2_multi.sh G'
G=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29511 \
  tests/multi_gpu_worker.py > gpurun_out/r2_multi_worker_$G.log 2>&1
echo "worker rc=$?" >> gpurun_out/r2_multi_worker_$G.log
tail -12 gpurun_out/r2_multi_worker_$G.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29512 \
  scripts/multi_gpu_timing.py > gpurun_out/r2_multi_timing_$G.log 2>&1
echo "timing rc=$?" >> gpurun_out/r2_multi_timing_$G.log
tail -9 gpurun_out/r2_multi_timing_$G.log
if [ "${2:-}" = "bench" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r2_bench_n$G.json 2> gpurun_out/r2_bench_n$G.err
  tail -c 400 gpurun_out/r2_bench_n$G.json
  python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_same_box.json 2>/dev/null
  tail -c 300 gpurun_out/r2_bench_n1_same_box.json
fi
