#!/bin/bash
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 60 python -m pytest tests/test_golden.py -m gpu -q 2>&1 | tail -2
