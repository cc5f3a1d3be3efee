#!/bin/bash
# A/B of the release side of the packed particle exchange (LLPF_XCHG_VARIANT 1 / 2 / 3) on 2 GPUs: parity worker + us per step
mkdir -p gpurun_out
O=gpurun_out/r2_xchg_ab.log; : > $O
V=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants
for v in 3 1 2; do
  if [ $v = 3 ]; then unset LLPF_LIB_PATH; else export LLPF_LIB_PATH=$V/libllpf_x$v.so; fi
  echo "== variant $v" | tee -a $O
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 2953$v \
    tests/multi_gpu_worker.py > gpurun_out/r2_xchg_worker_$v.log 2>&1; echo "worker rc=$? $(grep -c MULTI_GPU_OK gpurun_out/r2_xchg_worker_$v.log)" | tee -a $O
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 2954$v \
    scripts/multi_gpu_timing.py 2>&1 | grep "us/step" | tee -a $O
done
