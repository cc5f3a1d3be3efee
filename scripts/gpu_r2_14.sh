#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_enkf.py -m gpu -q > gpurun_out/r2n_pytest_enkf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_enkf.log
tail -60 gpurun_out/r2n_pytest_enkf.log
