"""One launch of the wide engine (config 5 shape) or of the smoother kernel, for ncu -c 1."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llpf_b200 as L
from llpf_b200 import workloads as W
what = sys.argv[1]
if what == "wide":
    s = W.lg_large_spec(seed=0); T, N = 20, 1 << 20
    u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(s, u, seed=1)
    pf = s.particle_filter(N, seed=1)
    print("ll", L.loglik(pf, u, y, epoch=1), "ms", L.last_run_ms(pf))
else:
    s = W.lg_spec(4, 2, 2, seed=0); T, N, M = 50, 1 << 16, 148
    u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(s, u, seed=1)
    pf = s.particle_filter(N, seed=2)
    xb, ll = L.smooth(pf, M, u, y, epoch=1)
    print("ll", ll, "smooth ms", L.last_smooth_ms(pf))
