"""One loglik launch for ncu: python scripts/prof_case.py <log2n> <T> <thr> [aux]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llpf_b200 as L
from llpf_b200 import workloads as W
log2n, T, thr = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
spec = W.lg_spec(4, 2, 2, seed=0)
u = np.random.default_rng(0).standard_normal((T, 2))
_, y = W.simulate_lg(spec, u, seed=1)
pf = (spec.aux_filter if len(sys.argv) > 4 else spec.particle_filter)(1 << log2n, seed=1, resample_threshold=thr)
d = L.loglik(pf, u, y, epoch=1, details=True)
print("ms", L.last_run_ms(pf), "rho", d["resampled"].mean(), "ll", d["ll"])
