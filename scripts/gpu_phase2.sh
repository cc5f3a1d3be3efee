#!/bin/bash
mkdir -p gpurun_out
{ echo "== 1 GPU"; timeout 200 python scripts/phase_timing.py 20 200 1.0; timeout 200 python scripts/phase_timing.py 20 200 0.0
echo "== 2 GPUs"; for thr in 1.0 0.0; do timeout 300 python -m torch.distributed.run --standalone --nnodes=1 --nproc-per-node 2 scripts/phase_timing_multi.py 20 200 $thr 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM"; done; } | tee gpurun_out/phases_multi.log
