#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_gpu.log
tail -6 gpurun_out/r2h_pytest_gpu.log
python scripts/tune.py quick > gpurun_out/r2h_tune_quick.log 2>&1; cat gpurun_out/r2h_tune_quick.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_c2.json 2>/dev/null
python -c "import json;d=json.loads(open('gpurun_out/r2h_bench_c2.json').read().strip().splitlines()[-1]);print('config 2: ms',d['ms_per_step'],'value %.3e'%d['value'],'frac',d['roofline']['frac'])"
