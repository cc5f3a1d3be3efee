#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_engineILi4 -c 1 --launch-skip 3 -o gpurun_out/r2_engine -f \
  python bench.py --T 200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_engine.log 2>&1
ls -la gpurun_out/r2_engine.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2_launches.csv
