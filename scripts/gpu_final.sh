#!/bin/bash
# end-of-round evidence on one GPU: tests, smoke, bench (both arms), ncu launch list, all-configs timing
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench N=1"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.json
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== configs"; timeout 600 python scripts/configs_timing.py 2>&1 | tee gpurun_out/configs_timing.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches.csv
