#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== tune default"
timeout 600 python scripts/tune.py full 2>&1 | tee gpurun_out/tune_v4.log
for v in b256_m3 b512_m1; do
  echo "== tune $v"
  LLPF_LIB_PATH=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants/libllpf_$v.so timeout 300 python scripts/tune.py quick 2>&1 | tee -a gpurun_out/tune_v4_variants.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_engine -c 1 -f -o gpurun_out/prof_v4_nores python scripts/prof_case.py 20 100 0.0 > gpurun_out/prof_v4_nores.log 2>&1
