"""Weak-scaling timing of the sharded filter: torchrun --standalone --nproc-per-node G scripts/multi_gpu_timing.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llpf_b200 as L
from llpf_b200 import workloads as W
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
spec = W.lg_spec(4, 2, 2, seed=0)
wide = W.lg_large_spec(seed=0)   # config 5: 64 states, Float32 particles
# per-GPU sizes; config 4 = APF with N=2^22 over 8 GPUs (2^19 each), config 5 = wide PF with N=2^20 over 8 GPUs (2^17 each)
for (log2n, T, thr, kind) in [(20, 300, 0.0, "pf"), (20, 300, 0.1, "pf"), (20, 300, 1.0, "pf"), (12, 300, 0.1, "pf"),
                              (19, 300, 0.1, "aux"), (17, 60, 0.1, "wide"), (20, 20, 0.1, "wide")]:
    sp = wide if kind == "wide" else spec
    u = np.random.default_rng(0).standard_normal((T, 2))
    _, y = W.simulate_lg(sp, u, seed=1)
    N = (1 << log2n) * world
    mk = sp.aux_filter if kind == "aux" else sp.particle_filter
    pf = mk(N, seed=1, resample_threshold=thr, device=local, rank=rank, world=world)
    L.connect_shards(pf)
    best = 1e9
    for rep in range(3):
        dist.barrier()
        r = L.loglik(pf, u, y, epoch=rep + 1, details=True)
        best = min(best, L.last_run_ms(pf))
    t = torch.tensor([best]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{kind} world={world} N={world}x2^{log2n} T={T} thr={thr}: {t.item():8.3f} ms {t.item()/T*1e3:7.2f} us/step "
              f"{N*T/t.item()/1e6:8.1f} Gps/s rho={r['resampled'].mean():.2f} ll={r['ll']:.4f}", flush=True)
dist.destroy_process_group()
