#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_estimation.py -m gpu -q 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_engine_wide -c 1 -o gpurun_out/r2_wide_v2 -f \
  python bench.py --config 5 --T 6 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_wide.log 2>&1
ls -la gpurun_out/r2_wide_v2.ncu-rep
