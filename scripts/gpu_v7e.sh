#!/bin/bash
V=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants
for v in mb3 mb4; do for z in 0 16; do echo "== $v zcap $z"; LLPF_LIB_PATH=$V/libllpf_$v.so LLPF_ZCAP=$z timeout 300 python scripts/tune.py quick 2>&1 | grep -v lib: ; done; done | tee gpurun_out/tune_v7e.log
