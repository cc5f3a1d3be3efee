#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_gpu.log
tail -30 gpurun_out/r2e_pytest_gpu.log
python scripts/tune.py residual > gpurun_out/r2e_tune_strategies.log 2>&1; cat gpurun_out/r2e_tune_strategies.log
