#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
tail -25 gpurun_out/r2b_pytest_gpu.log
python scripts/tune.py quick > gpurun_out/r2b_tune_quick.log 2>&1; cat gpurun_out/r2b_tune_quick.log
