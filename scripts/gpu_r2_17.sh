#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
python -c "import json;d=json.loads(open('gpurun_out/r2_bench_c2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'])"
