#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_estimation.py tests/test_gpu_parity.py -m gpu -q -k "statistics or smoother or callbacks or metropolis" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
grep -v "^$" gpurun_out/r2g_pytest.log | tail -40
