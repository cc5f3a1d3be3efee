#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_gpu.log
tail -6 gpurun_out/r2k_pytest_gpu.log
run() { # label, env, config
  LLPF_ZC="$2" python bench.py --config $3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2k_zc_tmp.json
  python -c "import json;d=json.loads(open('gpurun_out/r2k_zc_tmp.json').read());print('config $3 LLPF_ZC=$2 : ms %.3f  frac %.3f'%(d['ms_per_step'], d['roofline']['frac']))" | tee -a gpurun_out/r2k_zc_sweep.log
}
: > gpurun_out/r2k_zc_sweep.log
for z in 0 6,3,5,6 6,2,4,6 6,4,5,6 6,6,6,6 6,2,3,4 4,2,3,4 3,3,3,3 6,3,4,5; do run zc "$z" 2; done
for z in 0 6,3,5,6 6,6,6,6; do run zc "$z" 4; done
for z in 0 6,3,5,6; do run zc "$z" 3; done
