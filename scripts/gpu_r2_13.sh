#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rbpf.py tests/test_user_model.py -m gpu -q > gpurun_out/r2m_pytest_rbpf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_rbpf.log
tail -40 gpurun_out/r2m_pytest_rbpf.log
