"""Per-phase cycle breakdown of the engine (block 0's timeline) — needs the -DLLPF_PHASE_TIMING build."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LLPF_LIB_PATH"] = os.path.join(ROOT, "lowlevelparticlefilters.jl_b200/csrc/variants/libllpf_timing.so")
os.environ["LLPF_PHASE_DUMP"] = "/tmp/phases.bin"
import llpf_b200 as L
from llpf_b200 import workloads as W
log2n, T, thr = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
spec = W.lg_spec(4, 2, 2, seed=0)
u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(spec, u, seed=1)
pf = spec.particle_filter(1 << log2n, seed=1, resample_threshold=thr)
for rep in range(2):
    d = L.loglik(pf, u, y, epoch=rep + 1, details=True)
ts = np.fromfile("/tmp/phases.bin", dtype=np.int64).reshape(-1, 16)[1:T]  # passes k=1..T-1
res = d["resampled"][:T - 1].astype(bool)
ghz = 1.95
def seg(a, b, mask): 
    v = (ts[mask, b] - ts[mask, a]); v = v[(ts[mask, a] > 0) & (ts[mask, b] > 0)]
    return v.mean() / ghz / 1e3 if v.size else float('nan')
print(f"N=2^{log2n} thr={thr} ms={L.last_run_ms(pf):.3f} rho={res.mean():.2f}  (us, block 0)")
for name, m in (("non-resample", ~res), ("resample", res)):
    if m.sum() == 0: continue
    print(f" {name:13s} n={m.sum():4d}: scan1 {seg(0,1,m):6.2f} | bar {seg(1,2,m):6.2f} | offsets+scatter {seg(2,3,m):6.2f} | bar {seg(3,4,m):6.2f} | "
          f"main(from 0 or 4) {seg(4,5,m) if name=='resample' else seg(0,5,m):6.2f} | blk-reduce {seg(5,6,m):6.2f} | bar {seg(6,7,m):6.2f} | combine {seg(7,8,m):6.2f} | total {seg(0,8,m):6.2f}")
