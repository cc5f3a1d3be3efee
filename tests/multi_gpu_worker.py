"""Sharded-filter check, run with:  torchrun --standalone --nproc-per-node G tests/multi_gpu_worker.py
Every rank builds its shard of ONE global filter (particles block-partitioned over the GPUs), the kernels
exchange statistics / CDF offsets / offspring indices / resampled particles over NVLink peer memory, and the
result is compared with (a) the same filter on one GPU and (b) the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import llpf_b200 as L  # noqa: E402
from models import lg_large_model, lg_model  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # only carries the IPC descriptors; the data path is peer memory
    s4 = lg_model(4, 2, 2, seed=0)
    s64 = lg_large_model(seed=1)      # config 5: 64 states, Float32 particles; weights degenerate -> heavy offspring runs
    ok = True
    for (N, T, thr, kind) in [(4096, 60, 0.1, "pf"), (1 << 16, 80, 0.5, "pf"), (1 << 16, 40, 1.0, "pf"),
                              (1 << 14, 50, 0.1, "aux"), (1 << 20, 50, 0.1, "pf"), (1 << 13, 10, 0.5, "wide"),
                              (1 << 18, 6, 0.5, "wide")]:
        s = s64 if kind == "wide" else s4
        u = np.random.default_rng(3).standard_normal((T, 2))
        gen = s.oracle_filter(64, seed=1)
        _, y = gen.simulate(u, 17)
        mk = s.aux_filter if kind == "aux" else s.particle_filter
        pf = mk(N, seed=5, resample_threshold=thr, device=local, rank=rank, world=world)
        L.connect_shards(pf)
        r = L.loglik(pf, u, y, epoch=2, details=True)
        xs = L.particles(pf)
        ws = L.weights(pf)
        js = L.ancestors(pf)
        # sharded accessors (collectives: every rank calls them): the weights are normalised globally
        ess_s = L.effective_particles(pf)
        should_s = L.shouldresample(pf)
        xhat_s = L.weighted_mean(pf) if kind != "wide" else None
        gathered = [None] * world
        dist.all_gather_object(gathered, (r["ll"], xs, ws, js, r["resampled"]))
        if rank == 0:
            lls = [g[0] for g in gathered]
            assert all(v == lls[0] for v in lls), lls            # every rank computes the identical global ll
            X = np.concatenate([g[1] for g in gathered]); Wt = np.concatenate([g[2] for g in gathered])
            J = np.concatenate([g[3] for g in gathered])
            one = mk(N, seed=5, resample_threshold=thr, device=local)
            r1 = L.loglik(one, u, y, epoch=2, details=True)
            x1, w1, j1 = L.particles(one), L.weights(one), L.ancestors(one)
            ess1 = L.effective_particles(one)
            acc_ok = abs(ess_s - ess1) <= 1e-9 * ess1 and should_s == L.shouldresample(one)
            if xhat_s is not None:
                acc_ok &= bool(np.allclose(xhat_s, L.weighted_mean(one), rtol=1e-10, atol=1e-12))
            ok &= acc_ok
            rel = abs(lls[0] - r1["ll"]) / abs(r1["ll"])
            same_res = np.array_equal(r["resampled"], r1["resampled"])
            dx = np.abs(X - x1).max()
            nj = int((J != j1).sum())
            msg = f"{kind} N={N} T={T} thr={thr} world={world}: ll={lls[0]:.10f} vs 1-GPU {r1['ll']:.10f} rel={rel:.2e} " \
                  f"resampled_equal={same_res} max|dx|={dx:.2e} j_mismatch={nj} accessors_ok={acc_ok} ms={L.last_run_ms(pf):.2f} (1-GPU {L.last_run_ms(one):.2f})"
            if N <= (1 << 16) and kind != "wide" or N <= (1 << 13):
                ref = (s.oracle_filter(N, filter=2 if kind == "aux" else 0, seed=5, resample_threshold=thr)).loglik(u, y, epoch=2)
                relo = abs(lls[0] - ref["ll"]) / abs(ref["ll"])
                msg += f" | oracle rel={relo:.2e}"
                ok &= relo < 1e-6
            print(msg, flush=True)
            ok &= rel < 1e-9 and same_res and (dx < 1e-9 or nj > 0) and np.all(np.diff(J) >= 0 if r["resampled"][-1] else True)
            ok &= np.isfinite(Wt).all()
        dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
