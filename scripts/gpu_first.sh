#!/bin/bash
# first contact with the B200: sanity, sanitizer on the smoke case, GPU tests, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
echo "== sanitizer (memcheck) on smoke"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1
tail -15 gpurun_out/memcheck.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench"
timeout 600 python bench.py --steps 3 --warmup 2 2>&1 | tail -5 | tee gpurun_out/bench_first.log
