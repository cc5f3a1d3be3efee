"""Host-side mirror (numpy) of the sharded resampling protocol the engine runs on the device (SURVEY §8e,
csrc/llpf_engine.cuh: resample_indices / scatter_pairs / peer_allgather).  It documents the partition arithmetic
and is what the world_size-2 gloo tests drive on CPU; the product path itself runs entirely inside the CUDA kernels.

Particles are block-partitioned: rank r owns global indices [r*n, (r+1)*n), n = N/world.
  1. local inclusive scan of the weights in 2^-62 fixed point (uint64: exact, associative)
  2. all-gather of the ranks' totals -> this rank's global CDF offset (exact, so the global `bins` are
     bit-identical to a single-process scan, independent of `world`)
  3. source side: particle b owns the output slots {i : bins[b-1] <= s_i < bins[b]},  s_i = fl(r + fl(i*fl(1/N)))
     (src/resample.jl:23-34); F(v) = min{i : s_i >= v} is evaluated exactly
  4. slot i belongs to rank i // n.  Offspring that land on the source's own rank are scattered locally; a particle with
     slots on another rank is shipped there ONCE as a packed entry (state, first slot, count) and expanded by the
     destination (on the device: posted stores into the destination's `pack_in` over NVLink + counts on the cross-GPU
     barrier, csrc/llpf_engine.cuh push_remote_parts / peer_exchange_counts / expand_packs; here: an all-to-all).
This module is the CPU model of that protocol for the world_size-2/4 gloo tests; it is not on the product path.
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


# ---------------------------------------------------------------------------------------------
# Packed particle exchange (design for the next engine version; host mirror + gloo tests only)
# ---------------------------------------------------------------------------------------------
# The packed exchange (what the device does since round 2).  Round 1 routed offspring INDICES to the slot owners and every
# slot then gathered its ancestor's state from whichever rank held it (fine-grained peer loads: NVLink latency and
# same-address contention on heavy particles made resample steps 2-3x slower on 8 GPUs).  The source side already knows that
# particle b owns the slot range [F(lo_b), F(hi_b)), which is monotone in b, so for every destination rank d the particles
# with at least one slot on d are PACKED by the source (state + first slot + count), the destination expands the runs
# locally, and every needed particle crosses NVLink exactly once per destination rank, however many slots it owns there.
def pack_for_destinations(x_local, we_local, u01, N, rank, world, allgather):
    """Source side.  Returns, per destination rank d, (states [m_d, nx], first_slot_local [m_d], count [m_d]) for this
    rank's particles that own slots on d, plus (bins_local, f_total)."""
    first, n = shard_range(N, rank, world)
    f = to_fixed(we_local)
    loc = np.cumsum(f, dtype=np.uint64)
    tots = [int(t) for t in allgather(int(loc[-1]))]
    base, gtot = sum(tots[:rank]), sum(tots)
    inv = 2.0 ** -FIX_BITS
    hi = (np.uint64(base) + loc).astype(np.float64) * inv
    lo = np.concatenate([[float(base) * inv], hi[:-1]])
    total = float(gtot) * inv
    r = u01 * total / N
    fa, fc = first_slot_ge(lo, r, N), first_slot_ge(hi, r, N)      # slots [fa, fc) of every local particle
    f_total = int(first_slot_ge(np.array([total]), r, N)[0])
    packs = []
    for d in range(world):
        s0, s1 = d * n, (d + 1) * n
        a, c = np.maximum(fa, s0), np.minimum(fc, s1)               # intersection with d's slot range
        keep = c > a                                                # on the device: flags -> block scan -> pack index
        packs.append((np.asarray(x_local)[keep], (a[keep] - s0).astype(np.int64), (c[keep] - a[keep]).astype(np.int64)))
    return packs, hi, f_total


def expand_packs(recv, n, nx, x_prev_local=None):
    """Destination side: write every received particle into the run of local slots it owns.  Slots no pack covers are the
    reference's untouched entries (resample.jl:26-34): they keep the particle of the identity ancestor (x_prev_local)."""
    out = np.zeros((n, nx)) if x_prev_local is None else np.array(x_prev_local, dtype=np.float64).copy()
    covered = np.zeros(n, dtype=bool)
    moved = 0
    for states, first_slot, count in recv:
        moved += len(count)
        for st, a, c in zip(states, first_slot, count):
            out[a:a + c] = st
            covered[a:a + c] = True
    return out, covered, moved


def sharded_resample_packed(x_local, we_local, u01, N, rank, world, allgather, alltoall):
    """x_new_local[i] = x[j[i]] for this rank's slots without ever materialising remote gathers: pack -> exchange -> expand.
    Returns (x_new_local, covered mask, number of particle states received)."""
    first, n = shard_range(N, rank, world)
    packs, _, _ = pack_for_destinations(x_local, we_local, u01, N, rank, world, allgather)
    recv = alltoall(packs)
    return expand_packs(recv, n, np.asarray(x_local).shape[1], x_prev_local=x_local)

