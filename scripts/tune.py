"""Timing experiments on the GPU box: per-time-step cost of the engine under different regimes.
LLPF_LIB_PATH selects a tuning variant of the library."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llpf_b200 as L
from llpf_b200 import workloads as W


def run(log2n, T, thr, reps=3, filt="pf", **kw):
    spec = W.lg_spec(4, 2, 2, seed=0)
    u = np.random.default_rng(0).standard_normal((T, 2))
    _, y = W.simulate_lg(spec, u, seed=1)
    N = 1 << log2n
    pf = (spec.aux_filter if filt == "aux" else spec.particle_filter)(N, seed=1, resample_threshold=thr, **kw)
    best = 1e9
    for r in range(reps):
        d = L.loglik(pf, u, y, epoch=r + 1, details=True)
        best = min(best, L.last_run_ms(pf))
    rho = d["resampled"].mean()
    print(f"{filt} N=2^{log2n} T={T} thr={thr} {kw}: {best:8.3f} ms  {best / T * 1e3:7.2f} us/step  "
          f"{N * T / best / 1e6:9.1f} Gps/s  rho={rho:.3f} ll={d['ll']:.4f}", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "full"
    print("lib:", os.environ.get("LLPF_LIB_PATH", "default"), flush=True)
    if mode == "quick":
        for thr in (0.0, 0.1, 1.0):
            run(20, 300, thr)
        run(10, 300, 0.0)
    elif mode == "residual":
        for strat in (L.ResampleSystematic, L.ResampleStratified, L.ResampleResidual, L.ResampleMetropolis):
            print(strat.__name__, flush=True)
            for thr in (0.1, 1.0):
                run(20, 300, thr, resampling_strategy=strat)
            run(20, 300, 0.1, filt="aux", resampling_strategy=strat)
    else:
        for thr in (0.0, 0.1, 1.0):
            run(20, 300, thr)
        for n in (10, 14, 16, 18, 22):
            run(n, 300, 0.0)
            run(n, 300, 1.0)
        run(20, 300, 0.1, filt="aux")
        run(18, 50, 1.0, scan_mode="serial", reps=1)
