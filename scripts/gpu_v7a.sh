#!/bin/bash
# v7 step A: last-arriver combine + unswitched sweeps — parity, timing vs the variant without unswitching, phases
mkdir -p gpurun_out
V=$PWD/lowlevelparticlefilters.jl_b200/csrc/variants
echo "== tune default (la + unswitch)"; timeout 300 python scripts/tune.py quick 2>&1 | tee gpurun_out/tune_v7a.log
echo "== tune la only"; LLPF_LIB_PATH=$V/libllpf_la.so timeout 300 python scripts/tune.py quick 2>&1 | tee -a gpurun_out/tune_v7a.log
echo "== phases"; timeout 300 python scripts/phase_timing.py 20 300 0.1 2>&1 | tee gpurun_out/phases_v7a.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_v7a.json
