#!/bin/bash
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "reference_end_to_end" 2>&1 | tail -15
