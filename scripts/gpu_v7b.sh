#!/bin/bash
V=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants
export LLPF_LIB_PATH=$V/libllpf_na.so
for z in 16 8 4 0; do echo "== zcap $z"; LLPF_ZCAP=$z timeout 300 python scripts/tune.py quick 2>&1 | grep -v lib: ; done | tee gpurun_out/tune_v7b.log
unset LLPF_LIB_PATH
python scripts/skew.py 20 300 0.1 2>&1 | grep -v "=1\|=2" | tee gpurun_out/skew_v7b.log
