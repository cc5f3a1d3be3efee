#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_gpu.log
tail -30 gpurun_out/r2d_pytest_gpu.log
LLPF_LIB_PATH=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants/libllpf_b512.so LLPF_LIB_ALLOW_MISSING=1 python scripts/tune.py quick > gpurun_out/r2d_tune_b512.log 2>&1; cat gpurun_out/r2d_tune_b512.log
