#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rbpf.py -m gpu -q > gpurun_out/r2o_pytest_rbpf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest_rbpf.log
tail -5 gpurun_out/r2o_pytest_rbpf.log
