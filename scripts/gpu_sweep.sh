#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/tune_sweep.log
echo "== pytest gpu (default lib)"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for f in lowlevelparticlefilters.jl_b200/csrc/variants/*.so; do
  LLPF_LIB_PATH=$PWD/$f timeout 300 python scripts/tune.py quick 2>&1 | tee -a gpurun_out/tune_sweep.log
done
