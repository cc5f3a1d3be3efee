#!/bin/bash
# quick check: all GPU tests + the three tuning points + the three resampling strategies
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
{ echo "== new"; timeout 300 python scripts/tune.py quick 2>&1 | grep -v lib:
echo "== strategies"; timeout 300 python scripts/tune.py residual 2>&1 | grep -v lib: ; } | tee gpurun_out/tune_quick.log
