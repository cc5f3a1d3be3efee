"""Config 5 timing: ParticleFilter, 64-state LG, Float32 particles (llpf_wide.cuh).  us per time step and
particle-steps/s on one GPU; FP32-FMA roofline fraction (8192 + 64*ny FMAs per particle-step)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llpf_b200 as L
from llpf_b200 import workloads as W

def run(log2n, T, thr, reps=2, **kw):
    spec = W.lg_large_spec(seed=0)
    u = np.random.default_rng(0).standard_normal((T, 2))
    _, y = W.simulate_lg(spec, u, seed=1)
    N = 1 << log2n
    pf = spec.particle_filter(N, seed=1, resample_threshold=thr, **kw)
    best = 1e9
    for r in range(reps):
        d = L.loglik(pf, u, y, epoch=r + 1, details=True)
        best = min(best, L.last_run_ms(pf))
    fma = 64 * 64 + 64 * spec.ny + 64      # A x, G x', diag(L) z   (R1 = I)
    gps = N * T / best / 1e6
    print(f"wide N=2^{log2n} T={T} thr={thr}: {best:9.3f} ms {best / T * 1e3:9.2f} us/step {gps:7.2f} Gps/s  "
          f"{gps * fma * 2 / 1e3:6.1f} TFLOP/s fp32  rho={d['resampled'].mean():.2f} ll={d['ll']:.3f}", flush=True)

if __name__ == "__main__":
    for thr in (0.0, 0.1, 1.0):
        run(20, 30, thr)
    run(16, 30, 0.1)
    run(12, 30, 0.1)
