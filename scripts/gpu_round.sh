#!/bin/bash
# end-of-iteration evidence run: tests, bench (1 and 2 GPUs, both arms), ncu launch list + full capture
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench N=1"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n2.json
fi
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -4 gpurun_out/launches.csv
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_engine -c 1 -f -o gpurun_out/prof_full python bench.py --steps 1 --warmup 0 --T 200 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
