#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
for v in sc1 sc3 sc7; do
  echo "== tune $v"
  LLPF_LIB_PATH=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants/libllpf_$v.so timeout 300 python scripts/tune.py quick 2>&1 | tee -a gpurun_out/tune_v5_variants.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_engine -c 1 -f -o gpurun_out/prof_v5_nores python scripts/prof_case.py 20 100 0.0 > gpurun_out/prof_v5_nores.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_engine -c 1 -f -o gpurun_out/prof_v5_allres python scripts/prof_case.py 20 100 1.0 > gpurun_out/prof_v5_allres.log 2>&1
