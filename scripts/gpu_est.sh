#!/bin/bash
mkdir -p gpurun_out
echo "== pytest estimation"; timeout 600 python -m pytest tests/test_estimation.py -m gpu -q -x 2>&1 | tail -15
echo "== pmmh timing"; timeout 600 python scripts/pmmh_timing.py 2>&1 | tee gpurun_out/pmmh_timing.log
