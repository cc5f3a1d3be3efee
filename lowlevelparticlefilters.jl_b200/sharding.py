"""Host-side mirror (numpy) of the sharded resampling protocol the engine runs on the device (SURVEY §8e,
csrc/llpf_engine.cuh: resample_indices / scatter_pairs / peer_allgather).  It documents the partition arithmetic
and is what the world_size-2 gloo tests drive on CPU; the product path itself runs entirely inside the CUDA kernels.

Particles are block-partitioned: rank r owns global indices [r*n, (r+1)*n), n = N/world.
  1. local inclusive scan of the weights in 2^-62 fixed point (uint64: exact, associative)
  2. all-gather of the ranks' totals -> this rank's global CDF offset (exact, so the global `bins` are
     bit-identical to a single-process scan, independent of `world`)
  3. source side: particle b owns the output slots {i : bins[b-1] <= s_i < bins[b]},  s_i = fl(r + fl(i*fl(1/N)))
     (src/resample.jl:23-34); F(v) = min{i : s_i >= v} is evaluated exactly
  4. slot i belongs to rank i // n: the (slot, ancestor) pairs are routed to their owners (on the device: stores
     into the owner's `j` array over NVLink; here: an all-to-all)
"""
import numpy as np

FIX_BITS = 62


def shard_range(N, rank, world):
    if N % world:
        raise ValueError("N must be divisible by the number of ranks")
    n = N // world
    return rank * n, n


def to_fixed(we):
    v = np.asarray(we, dtype=np.float64) * float(1 << FIX_BITS)
    v = np.where(v > 0, np.minimum(v, float(1 << FIX_BITS)), 0.0)
    return np.rint(v).astype(np.uint64)


def thresholds(i, r, M):
    """s[i] = fl(r + fl(i * fl(1/M)))  — product and sum rounded separately (no FMA)."""
    return r + (np.asarray(i, dtype=np.float64) * (1.0 / M))


def first_slot_ge(v, r, M):
    """F(v) = min{ i in [0, M] : s_i >= v }  (vectorised; exact fix-up like the device code)."""
    v = np.asarray(v, dtype=np.float64)
    i0 = np.clip(np.ceil((v - r) * M), 0, M).astype(np.int64)
    for _ in range(4):
        down = (i0 > 0) & (thresholds(i0 - 1, r, M) >= v)
        i0 = np.where(down, i0 - 1, i0)
        up = (i0 < M) & (thresholds(np.minimum(i0, M - 1), r, M) < v)
        i0 = np.where(up, i0 + 1, i0)
    return i0


def local_offspring(we_local, u01, N, rank, world, allgather):
    """Steps 1-3 for one rank. Returns (slots, ancestors) as global 0-based arrays, plus (bins_local, f_total)."""
    first, n = shard_range(N, rank, world)
    f = to_fixed(we_local)
    loc = np.cumsum(f, dtype=np.uint64)
    tots = [int(t) for t in allgather(int(loc[-1]))]
    base = sum(tots[:rank])
    gtot = sum(tots)
    inv = 2.0 ** -FIX_BITS
    hi = (np.uint64(base) + loc).astype(np.float64) * inv
    lo = np.concatenate([[float(base) * inv], hi[:-1]])
    total = float(gtot) * inv
    r = u01 * total / N
    fa, fc = first_slot_ge(lo, r, N), first_slot_ge(hi, r, N)
    cnt = fc - fa
    anc = np.repeat(np.arange(first, first + n, dtype=np.int64), cnt)
    slots = np.concatenate([np.arange(a, c) for a, c in zip(fa, fc)]) if cnt.sum() else np.zeros(0, dtype=np.int64)
    f_total = int(first_slot_ge(np.array([total]), r, N)[0])
    return slots.astype(np.int64), anc, hi, f_total


def sharded_systematic(we_local, u01, N, rank, world, allgather, alltoall, j_prev=None):
    """Full protocol for one rank: returns this rank's slice of j (global 0-based ancestors) and its bins."""
    first, n = shard_range(N, rank, world)
    slots, anc, bins, f_total = local_offspring(we_local, u01, N, rank, world, allgather)
    dest = slots // n
    out = [(slots[dest == d] - d * n, anc[dest == d]) for d in range(world)]
    recv = alltoall(out)
    j = np.arange(first, first + n, dtype=np.int64) if j_prev is None else np.array(j_prev, dtype=np.int64)
    for s, a in recv:
        j[s] = a                                   # untouched (stale) slots keep their previous value
    return j, bins, f_total
