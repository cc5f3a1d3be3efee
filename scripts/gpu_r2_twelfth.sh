#!/bin/bash
mkdir -p gpurun_out
V=lowlevelparticlefilters.jl_b200/csrc/variants
log=gpurun_out/r2l_ab.log; : > $log
q() { echo "== $1 ZC=$2" | tee -a $log; LLPF_LIB_PATH=$3 LLPF_ZC="$2" python scripts/tune.py quick 2>&1 | grep "us/step" | tee -a $log; }
q prev - $PWD/$V/libllpf_prev.so
q new 0 ""
q new 6,2,4,6 ""
q new 6,3,5,6 ""
q new 6,2,3,5 ""
q new 6,1,3,5 ""
q new 6,2,6,6 ""
q prev - $PWD/$V/libllpf_prev.so
for z in prev 0 6,2,4,6; do
  for c in 4 3; do
    if [ $z = prev ]; then export LLPF_LIB_PATH=$PWD/$V/libllpf_prev.so; else unset LLPF_LIB_PATH; fi
    LLPF_ZC="$z" python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2l_tmp.json
    python -c "import json;d=json.loads(open('gpurun_out/r2l_tmp.json').read());print('config $c $z : ms %.3f'%(d['ms_per_step']))" | tee -a $log
  done
done
