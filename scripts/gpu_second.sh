#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== tune"
timeout 900 python scripts/tune.py 2>&1 | tee gpurun_out/tune.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches_r1.csv
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_engine -c 1 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 0 --T 200 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
