#!/bin/bash
# round 2, first GPU call: clean sm_100a build ON the GPU box, the whole GPU suite, bench lines of every BASELINE config
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt
( time python __graft_entry__.py --force ) > gpurun_out/r2_clean_build.log 2>&1
cat lowlevelparticlefilters.jl_b200/csrc/build_info.json >> gpurun_out/r2_clean_build.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
for c in 3 4 5; do python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err; done
tail -c 600 gpurun_out/r2_bench_c2.json
