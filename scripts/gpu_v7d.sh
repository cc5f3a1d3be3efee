#!/bin/bash
V=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants
export LLPF_LIB_PATH=$V/libllpf_ilp2.so
for z in 0 16; do echo "== ilp2 zcap $z"; LLPF_ZCAP=$z timeout 300 python scripts/tune.py quick 2>&1 | grep -v lib: ; done | tee gpurun_out/tune_v7d.log
unset LLPF_LIB_PATH
LLPF_ZCAP=0 python scripts/skew.py 20 300 0.0 2>&1 | grep -v "=1\|=2\|SM:mean" | tee gpurun_out/skew_v7d.log
