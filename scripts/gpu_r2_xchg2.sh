#!/bin/bash
# A/B, acquire side of the packed exchange: default (fence.acq_rel.sys) vs ld.acquire.sys with release variants 3 and 2
mkdir -p gpurun_out
O=gpurun_out/r2_xchg_ab2.log; : > $O
V=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants
for v in default y3 y2; do
  if [ $v = default ]; then unset LLPF_LIB_PATH; else export LLPF_LIB_PATH=$V/libllpf_$v.so; fi
  echo "== $v" | tee -a $O
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561 \
    tests/multi_gpu_worker.py > gpurun_out/r2_xchg2_worker_$v.log 2>&1; echo "worker rc=$? $(grep -c MULTI_GPU_OK gpurun_out/r2_xchg2_worker_$v.log)" | tee -a $O
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29562 \
    scripts/multi_gpu_timing.py 2>&1 | grep "us/step" | grep -v "2^12\|2^17" | tee -a $O
done
