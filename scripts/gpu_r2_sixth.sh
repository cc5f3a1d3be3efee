#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.log
tail -30 gpurun_out/r2f_pytest_gpu.log
python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err
python -c "import json;d=json.load(open('gpurun_out/r2f_bench_c5.json'));print('config 5: ms',d['ms_per_step'],'value %.3e'%d['value'],'fp32 frac',d['roofline_fp32']['frac'])"
