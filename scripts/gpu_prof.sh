#!/bin/bash
mkdir -p gpurun_out
for c in "20 100 0.0 nores" "20 100 1.0 allres"; do set -- $c
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_engine -c 1 -f -o gpurun_out/prof_v2_$4 python scripts/prof_case.py $1 $2 $3 > gpurun_out/prof_v2_$4.log 2>&1
tail -2 gpurun_out/prof_v2_$4.log
done
