#!/bin/bash
mkdir -p gpurun_out
for s in 0 -0.02 -0.04 -0.06 -0.09; do echo "== LLPF_WAVE_SKEW=$s"; LLPF_WAVE_SKEW=$s timeout 200 python scripts/tune.py quick 2>&1 | grep -v "lib:\|2^10"; done | tee gpurun_out/tune_wave_skew.log
