#!/bin/bash
mkdir -p gpurun_out
echo "== pytest smoother + golden"; timeout 300 python -m pytest tests -m gpu -q -k "smoother or golden" 2>&1 | tail -6
echo "== smoother timing"; timeout 300 python scripts/smooth_timing.py 2>&1 | tee gpurun_out/smooth_timing.log
