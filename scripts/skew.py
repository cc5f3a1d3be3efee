"""Per-block skew of the engine passes (needs the -DLLPF_PHASE_TIMING build): distribution over blocks of the
sweep-end time and of the time the finished statistics were seen, relative to the earliest block."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LLPF_LIB_PATH"] = os.path.join(ROOT, "lowlevelparticlefilters.jl_b200/csrc/variants/libllpf_timing.so")
os.environ["LLPF_PHASE_DUMP"] = "/tmp/phases.bin"
import llpf_b200 as L
from llpf_b200 import workloads as W
log2n, T, thr = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
MAXB = 1024
spec = W.lg_spec(4, 2, 2, seed=0)
u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(spec, u, seed=1)
pf = spec.particle_filter(1 << log2n, seed=1, resample_threshold=thr)
for rep in range(2):
    d = L.loglik(pf, u, y, epoch=rep + 1, details=True)
raw = np.fromfile("/tmp/phases.bin", dtype=np.int64)
blk = raw[16 * (T + 2):].reshape(T + 2, MAXB, 4)
nb = int((blk[5, :, 0] > 0).sum())
res = d["resampled"].astype(bool)
print(f"N=2^{log2n} thr={thr} ms={L.last_run_ms(pf):.3f} nblocks={nb}")
for name, m in (("non-resample", ~res), ("resample", res)):
    ks = [k for k in range(2, T - 1) if m[k - 1]]
    if not ks: continue
    b = blk[ks][:, :nb, :].astype(np.float64) / 1e3   # us
    t0 = b[:, :, 0].min(axis=1, keepdims=True)
    start, end, seen = b[:, :, 0] - t0, b[:, :, 1] - t0, b[:, :, 2] - t0
    scat = b[:, :, 3] - t0
    q = lambda a: "min %6.2f med %6.2f p90 %6.2f max %6.2f" % (a.min(axis=1).mean(), np.median(a, axis=1).mean(), np.percentile(a, 90, axis=1).mean(), a.max(axis=1).mean())
    print(f" {name} ({len(ks)} passes), us after the first block's pass start:")
    print("   pass start  :", q(start))
    if name == "resample": print("   indices done:", q(scat))
    print("   sweep end   :", q(end))
    print("   stats seen  :", q(seen))
    dur = end - (scat if name == "resample" else start)
    print("   sweep length:", q(dur), " slowest blocks:", np.argsort(-dur.mean(axis=0))[:8], " fastest:", np.argsort(dur.mean(axis=0))[:4])
    if name == "non-resample":
        smid = blk[0, :nb, 3]
        md = dur.mean(axis=0)
        order = np.argsort(md)
        print("   per-block mean sweep length (us), sorted; block:smid")
        for lo in range(0, nb, 37):
            print("    ", " ".join(f"{order[q]}:{smid[order[q]]}={md[order[q]]:.1f}" for q in range(lo, min(lo + 37, nb), 4)))
        # per-SM: both resident blocks
        bysm = {}
        for b in range(nb): bysm.setdefault(int(smid[b]), []).append((b, md[b]))
        sm_mean = sorted((np.mean([v for _, v in lst]), sm, lst) for sm, lst in bysm.items())
        pass
        pass
        print("   SM:mean sorted:", " ".join(f"{sm}:{m:.1f}" for m, sm, _ in sm_mean))
