#!/bin/bash
# residual resampling on the device: parity tests, A/B timing against the previous build, timing of the three strategies
mkdir -p gpurun_out
echo "== pytest residual"; timeout 600 python -m pytest tests -m gpu -q -x -k "residual or proportions" 2>&1 | tail -15
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
V=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants
{ for rep in 1 2; do
echo "== head"; LLPF_LIB_ALLOW_MISSING=1 LLPF_LIB_PATH=$V/libllpf_head.so timeout 300 python scripts/tune.py quick 2>&1 | grep -v lib:
echo "== new"; timeout 300 python scripts/tune.py quick 2>&1 | grep -v lib:
done
echo "== strategies"; timeout 300 python scripts/tune.py residual 2>&1 | grep -v lib: ; } | tee gpurun_out/tune_v8_residual.log
