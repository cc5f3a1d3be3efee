#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "wide or f32 or golden" > gpurun_out/r2j_pytest_wide.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_wide.log
tail -15 gpurun_out/r2j_pytest_wide.log
python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_c5.json 2> gpurun_out/r2j_bench_c5.err
python -c "import json;d=json.loads(open('gpurun_out/r2j_bench_c5.json').read().strip().splitlines()[-1]);print('config 5: ms',d['ms_per_step'],'value %.3e'%d['value'],'fp32 frac',d['roofline_fp32']['frac'])"
