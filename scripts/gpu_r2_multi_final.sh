#!/bin/bash
# round 2, final binary on an 8-GPU box: the multi-GPU parity test inside pytest (torchrun 2 and 8), per-step timing, weak scaling
G=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2_gpus_multi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sharded" -s > $O/r2_pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest_multi.log
grep -E "MULTI_GPU_OK|passed|failed|skipped" $O/r2_pytest_multi.log | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29512 \
  scripts/multi_gpu_timing.py > $O/r2_multi_timing_$G.log 2>&1
echo "timing rc=$?" >> $O/r2_multi_timing_$G.log
tail -9 $O/r2_multi_timing_$G.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $G --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_n$G.json 2> $O/r2_bench_n$G.err
tail -c 500 $O/r2_bench_n$G.json
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_n1_8gpubox.json 2>/dev/null
tail -c 300 $O/r2_bench_n1_8gpubox.json
