#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_user_model.py tests/test_estimation.py tests/test_golden.py "tests/test_gpu_parity.py::test_headline_size_config2_against_oracle" "tests/test_gpu_parity.py::test_systematic_follows_julia_rational_range" -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -40 gpurun_out/r2c_pytest.log
timeout 600 python scripts/pmmh_timing.py > gpurun_out/r2c_pmmh_timing.log 2>&1; cat gpurun_out/r2c_pmmh_timing.log
