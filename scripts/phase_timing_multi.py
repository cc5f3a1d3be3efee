import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
os.environ["LLPF_LIB_PATH"] = os.path.join(ROOT, "lowlevelparticlefilters.jl_b200/csrc/variants/libllpf_timing.so")
os.environ["LLPF_PHASE_DUMP"] = f"/tmp/phases_{rank}.bin"
import llpf_b200 as L
from llpf_b200 import workloads as W
torch.cuda.set_device(rank); dist.init_process_group("gloo")
log2n, T, thr = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
spec = W.lg_spec(4, 2, 2, seed=0)
u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(spec, u, seed=1)
pf = spec.particle_filter((1 << log2n) * world, seed=1, resample_threshold=thr, device=rank, rank=rank, world=world)
L.connect_shards(pf)
for rep in range(2):
    dist.barrier(); d = L.loglik(pf, u, y, epoch=rep + 1, details=True)
if rank == 0:
    ts = np.fromfile("/tmp/phases_0.bin", dtype=np.int64).reshape(-1, 16)[1:T]
    res = d["resampled"][:T - 1].astype(bool); ghz = 1.95
    def seg(a, b, m):
        v = ts[m, b] - ts[m, a]; v = v[(ts[m, a] > 0) & (ts[m, b] > 0)]; return v.mean() / ghz / 1e3 if v.size else float("nan")
    print(f"world={world} N=2^{log2n}/gpu thr={thr} ms={L.last_run_ms(pf):.3f}")
    for name, m in (("non-res", ~res), ("res", res)):
        if m.sum() == 0: continue
        if name == "res":
            print(f"   detail : offsets+totals-xchg {seg(2,14,m):6.2f} | scatter+push {seg(14,3,m):6.2f} | fence+bar {seg(3,10,m):6.2f} | "
                  f"counts-xchg {seg(10,11,m):6.2f} | expand {seg(11,12,m):6.2f} | bar {seg(12,13,m):6.2f} | heavy-fill {seg(13,4,m):6.2f}")
        print(f" {name:8s}: scan1 {seg(0,1,m):6.2f} | bar {seg(1,2,m):6.2f} | scatter(+totals xchg) {seg(2,3,m):6.2f} | bar+peerbar {seg(3,4,m):6.2f} | main {seg(4,5,m) if name=='res' else seg(0,5,m):6.2f} | "
              f"blk-red {seg(5,6,m):6.2f} | bar {seg(6,7,m):6.2f} | combine {seg(7,9,m):6.2f} | xchg {seg(9,8,m):6.2f} | total {seg(0,8,m):6.2f}")
dist.barrier()
if rank == 1:   # rank 1's thread 0 is the one that polls rank 0's count word: where the counts exchange spends its time
    ts = np.fromfile("/tmp/phases_1.bin", dtype=np.int64).reshape(-1, 16)[1:T]
    res = d["resampled"][:T - 1].astype(bool); ghz = 1.95
    def seg1(a, b, m):
        v = ts[m, b] - ts[m, a]; v = v[(ts[m, a] > 0) & (ts[m, b] > 0)]; return v.mean() / ghz / 1e3 if v.size else float("nan")
    print(f"   rank 1 : fence+bar {seg1(3,10,res):6.2f} | counts: post+poll {seg1(10,15,res):6.2f} | fence.sys after the poll {seg1(15,11,res):6.2f} | "
          f"expand {seg1(11,12,res):6.2f} | bar {seg1(12,13,res):6.2f}")
dist.destroy_process_group()
