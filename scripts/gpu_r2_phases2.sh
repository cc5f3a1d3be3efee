#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2_phases_2_after.log; : > $O
for thr in 1.0 0.1; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29521 \
    scripts/phase_timing_multi.py 20 200 $thr >> $O 2>&1
done
# one GPU, same build, for the per-phase difference
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=1 --master-addr 127.0.0.1 --master-port 29522 \
  scripts/phase_timing_multi.py 20 200 1.0 >> $O 2>&1
grep -v "^\[\|^W\|^\*" $O | tail -30
