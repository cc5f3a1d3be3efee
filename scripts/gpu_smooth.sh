#!/bin/bash
mkdir -p gpurun_out
echo "== pytest smoother"; timeout 600 python -m pytest tests -m gpu -q -x -k "smoother" 2>&1 | tail -25
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== smoother timing"; timeout 600 python scripts/smooth_timing.py 2>&1 | tee gpurun_out/smooth_timing.log
