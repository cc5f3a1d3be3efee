#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_smooth -c 1 -f -o gpurun_out/prof_smooth python scripts/prof_extra.py smooth 2>&1 | tail -2
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_engine_wide -c 1 -f -o gpurun_out/prof_wide python scripts/prof_extra.py wide 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
