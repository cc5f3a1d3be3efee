"""FFBS smoother timing on the GPU box: backward-simulation kernel time (CUDA events) for a few (N, T, M), next to the
CPU oracle's backward pass on a bounded sample (smaller T), both as pair evaluations (M*N*(T-1)) per second."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llpf_b200 as L
from models import lg_model

s = lg_model(4, 2, 2, seed=0)
for (N, T, M) in [(2000, 200, 100), (1 << 14, 200, 128), (1 << 16, 200, 148), (1 << 16, 100, 1184), (1 << 18, 50, 296), (1 << 20, 20, 296)]:
    u = np.random.default_rng(0).standard_normal((T, 2))
    gen = s.oracle_filter(64, seed=1)
    _, y = gen.simulate(u, 3)
    pf = s.particle_filter(N, seed=2)
    best = 1e9
    for rep in range(2):
        t0 = time.perf_counter()
        xb, ll = L.smooth(pf, M, u, y, epoch=1)
        wall = time.perf_counter() - t0
        best = min(best, L.last_smooth_ms(pf))
    pairs = M * N * (T - 1)
    print(f"N={N} T={T} M={M}: backward kernel {best:9.3f} ms  {pairs / best / 1e6:8.2f} G pair-evals/s  "
          f"(forward {L.last_run_ms(pf):.2f} ms, call wall {wall * 1e3:.1f} ms) ll={ll:.4f}", flush=True)
# CPU oracle on a bounded sample
N, T, M = 2000, 40, 20
u = np.random.default_rng(0).standard_normal((T, 2))
of = s.oracle_filter(N, seed=2)
_, y = of.simulate(u, 3)
sol = of.forward_trajectory(u, y, epoch=1, history=True)
t0 = time.perf_counter()
of.smooth(M, u, sol["x"], sol["w"], sol["we"], epoch=1)
dt = time.perf_counter() - t0
print(f"CPU oracle (1 core) N={N} T={T} M={M}: {dt * 1e3:.1f} ms  {M * N * (T - 1) / dt / 1e9:.4f} G pair-evals/s")
