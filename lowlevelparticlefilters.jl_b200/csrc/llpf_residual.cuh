// llpf_residual.cuh — resample(ResampleResidual, we, j, bins, M) on the device  (reference src/resample.jl:63-117).
//
// The reference: wsum = sum(we) ; nw_i = we_i * (1/wsum) * M ; cnt_i = floor(nw_i) copies of i into the first
// `num = sum(cnt)` slots ; bins_i = nw_i - cnt_i ; if num < M: bins <- cumsum(bins / sum(bins)) and every remaining
// slot m draws u = rand() and takes the first i with u < bins[i] (a slot whose u is not below bins[N] keeps its old
// content).  Included from llpf_engine.cuh (inside no namespace; uses the scan / scatter helpers defined there).
//
// Device formulation (one cooperative grid, 3-4 grid barriers, no dependent chain except the per-draw bisection):
//   A  we_i -> bins[i] (stash) ; wsum: FAST = exact fixed-point grid sum, SERIAL = one thread, left-to-right f64
//   B  cnt_i / residual_i ; block-local scans of both (integers; residuals in fixed point) ; block totals
//   C  block offsets ; deterministic part from the SOURCE side (particle i owns slots [C_{i-1}, C_i), scatter_runs) ;
//      bins[i] = normalised residual CDF (FAST: exact integer prefix / total, SERIAL: the reference's three serial loops)
//   D  every block fills its own output slots >= num: u_k (k = m - num, in the reference's draw order) -> upper-bound
//      bisection in bins.
// SERIAL mode reproduces the reference bit for bit (same operation order for wsum, rsum, the scaling and the cumsum);
// FAST mode is bit-identical whenever the sums are exactly representable and otherwise differs like any re-associated sum.
#pragma once

namespace llpf {

// where the normalised weights come from (no lambdas: the routine is deliberately NOT inlined into the sweeps)
struct WeSrc {
  const double* w;   // mode 0: plain weights ; 2: log-weights, we = exp(w) ; 3: raw log-weights, we = exp(w - pm) * inv_s
  int mode;          // 1: uniform, we = weu
  double pm, pls, inv_s, weu, wu;
  const MathTab* T;
  double* hist_w;    // optional history rows (already offset to this step / shard): normalised log-weights, weights
  double* hist_we;
};
__device__ __forceinline__ double wesrc_we(const WeSrc& s, int i) {
  if (s.mode == 1) return s.weu;
  const double wr = __ldcg(s.w + i);
  if (s.mode == 0) return wr;
  if (s.mode == 2) return exp(wr);
  return exp_nonpos(wr - s.pm, *s.T) * s.inv_s;
}
__device__ __forceinline__ double wesrc_wnorm(const WeSrc& s, int i) {
  if (s.mode == 1) return s.wu;
  const double wr = __ldcg(s.w + i);
  return (s.mode == 3) ? (wr - s.pm) - s.pls : wr;
}

__device__ __forceinline__ u64 block_sum_u64(u64 v, Shared& sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.wtot[threadIdx.x >> 5] = v;
  __syncthreads();
  u64 t = 0;
#pragma unroll
  for (int k = 0; k < NWARP; ++k) t += sh.wtot[k];
  return t;
}

// u_draws: the rand() values of resample.jl:106 in draw order (stand-alone entry), nullptr -> counter RNG (ST_RESID).
// jid: state.j is logically 1:N (filtering.jl:148): a slot that keeps "its old content" gets its own index.
// Afterwards j[] is valid on [slot_lo, slot_hi) for the calling block and everywhere after the next grid barrier.
template <class JT>
__device__ __noinline__ void resample_residual(const EngineP& P, Shared& sh, int beg, int end, unsigned& bar_target,
                                               const WeSrc src, const double* u_draws, uint32_t step_idx, int Mslots,
                                               JT* jout_flat, JT jbase, int jid, int slot_lo, int slot_hi,
                                               double& total_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool fast = (P.scan_mode == 0);
  if (P.heavy != nullptr && LLPF_BLOCKIDX == 0 && threadIdx.x == 0) __stcg(P.heavy, 0);
  // the block partials of the statistics reduction are free between two reductions: pass A's block totals go there
  // (tots / tots2 are written by pass B while slower blocks may still be summing pass A's totals)
  u64* const pa = reinterpret_cast<u64*>(P.partials);
  // ---- A: stash the weights, total weight --------------------------------------------------------------------
  {
    u64 fsum = 0;
    for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
      const double we = wesrc_we(src, i);
      if (src.hist_w) {
        __stcs(src.hist_w + i, wesrc_wnorm(src, i));
        __stcs(src.hist_we + i, we);
      }
      __stcg(P.bins + i, we);
      if (fast) fsum += to_fixed(we, P.fix_scale);
    }
    if (fast) {
      const u64 t = block_sum_u64(fsum, sh);
      if (threadIdx.x == 0) __stcg(pa + LLPF_BLOCKIDX, t);
    }
  }
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  double wsum;
  if (fast) {
    scan_block_offsets_from(pa, P.nblocks, sh, 0ull);
    wsum = (double)sh.offs[P.nblocks] * P.fix_inv;
    __syncthreads();
  } else {
    if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0) {   // resample.jl:66-69
      double acc = 0.0;
      for (int i = 0; i < P.n; ++i) acc = __dadd_rn(acc, __ldcg(P.bins + i));
      __stcg(P.partials + MAX_BLOCKS, acc);
    }
    grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
    wsum = __ldcg(P.partials + MAX_BLOCKS);
  }
  const double inv_wsum = __ddiv_rn(1.0, wsum);            // :71
  const double Md = (double)Mslots;
  // residuals < 1 and there are at most M of them: 2^(62 - bits(M)) keeps their fixed-point sum below 2^62
  const int rbits = 62 - (32 - __clz(Mslots));
  const double rscale = __longlong_as_double((long long)(1023 + rbits) << 52);
  const double rinv = __longlong_as_double((long long)(1023 - rbits) << 52);
  // ---- B: counts, residuals, block-local scans ------------------------------------------------------------------
  {
    u64 carry_r = 0;
    unsigned carry_c = 0;
    for (int base = beg; base < end; base += BLOCK) {
      const int i = base + threadIdx.x;
      unsigned cnt = 0;
      double resid = 0.0;
      if (i < end) {
        const double nw = __dmul_rn(__dmul_rn(__ldcg(P.bins + i), inv_wsum), Md);   // :76
        if (nw >= 1.0) cnt = (unsigned)__double2ll_rd(fmin(nw, 2147483647.0));      // :77 floor
        resid = nw - (double)cnt;                                                   // :78 (exact)
      }
      u64 rf = fast ? to_fixed(resid, rscale) : 0ull;
      unsigned c = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 tr = __shfl_up_sync(0xffffffffu, rf, o);
        const unsigned tc = __shfl_up_sync(0xffffffffu, c, o);
        if (lane >= o) { rf += tr; c += tc; }
      }
      __syncthreads();
      if (lane == 31) { sh.wtot[warp] = rf; sh.wtf[warp] = (int)c; }
      __syncthreads();
      u64 woff_r = 0, rtot = 0;
      unsigned woff_c = 0, ctot = 0;
#pragma unroll
      for (int k = 0; k < NWARP; ++k) {
        const u64 tr = sh.wtot[k];
        const unsigned tc = (unsigned)sh.wtf[k];
        if (k < warp) { woff_r += tr; woff_c += tc; }
        rtot += tr; ctot += tc;
      }
      if (i < end) {
        const u64 packed = ((u64)(carry_c + woff_c + c) << 32) | (u64)cnt;   // (block-local inclusive count, own count)
        if (fast) {
          __stcg(P.loc + i, carry_r + woff_r + rf);
          __stcg(reinterpret_cast<u64*>(P.bins) + i, packed);
        } else {
          __stcg(P.loc + i, packed);
          __stcg(P.bins + i, resid);
        }
      }
      carry_r += rtot;
      carry_c += ctot;
    }
    if (threadIdx.x == 0) {
      __stcg(P.tots + LLPF_BLOCKIDX, carry_r);
      __stcg(P.tots2 + LLPF_BLOCKIDX, (u64)carry_c);
    }
  }
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  // ---- C: offsets, deterministic copies, residual CDF ------------------------------------------------------------
  scan_block_offsets_from(P.tots2, P.nblocks, sh, 0ull);
  const int coff = (int)sh.offs[LLPF_BLOCKIDX];
  const int num = (int)sh.offs[P.nblocks];
  __syncthreads();
  u64 roff = 0, rtot_all = 0;
  if (fast) {
    scan_block_offsets_from(P.tots, P.nblocks, sh, 0ull);
    roff = sh.offs[LLPF_BLOCKIDX];
    rtot_all = sh.offs[P.nblocks];
    __syncthreads();
  } else if (num != Mslots) {
    if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0) {
      double rsum = 0.0;                                              // :89-92
      for (int i = 0; i < P.n; ++i) rsum = __dadd_rn(rsum, __ldcg(P.bins + i));
      const double inv_rsum = __ddiv_rn(1.0, rsum);                   // :94
      double acc = __dmul_rn(__ldcg(P.bins), inv_rsum);               // :95-97, :99-102
      __stcg(P.bins, acc);
      for (int i = 1; i < P.n; ++i) {
        acc = __dadd_rn(__dmul_rn(__ldcg(P.bins + i), inv_rsum), acc);
        __stcg(P.bins + i, acc);
      }
    }
    __syncthreads();
  }
  SlotRouter<JT> jout;
  jout.heavy = P.heavy;
  jout.j = jout_flat;
  jout.base = P.first;
  const double rden = (double)rtot_all;
  for (int base = beg; base < end; base += BLOCK) {   // uniform trip count: warp collectives inside
    const int i = base + threadIdx.x;
    int lo = 0, cnt = 0;
    if (i < end) {
      const u64 packed = fast ? __ldcg(reinterpret_cast<const u64*>(P.bins) + i) : __ldcg(P.loc + i);
      cnt = (int)(packed & 0xffffffffull);
      lo = coff + (int)(packed >> 32) - cnt;
      if (fast) {
        const u64 pr = __ldcg(P.loc + i);
        double b;
        if (num != Mslots) {
          b = __ddiv_rn((double)(roff + pr), rden);                   // exact integer prefix, one rounding
        } else {                                                      // :85-87 early return: bins holds the residuals
          const u64 prev = (i > beg) ? __ldcg(P.loc + i - 1) : 0ull;
          b = (double)(pr - prev) * rinv;
        }
        __stcg(P.bins + i, b);
      }
    }
    scatter_runs<JT>(jout, lo, cnt, jbase + (JT)i);
  }
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  // ---- D: heavy runs, multinomial part ----------------------------------------------------------------------------
  double hcnt[MAX_WORLD][1];
  hcnt[0][0] = (P.heavy != nullptr) ? (double)__ldcg(P.heavy) : 0.0;
  {
    bool any = hcnt[0][0] > 0.0;
    if (any) {
      int n = (int)hcnt[0][0];
      if (n > HEAVY_MAX) n = HEAVY_MAX;
      for (int e = 0; e < n; ++e) {
        const int lo = __ldcg(P.heavy + 1 + 3 * e), c = __ldcg(P.heavy + 2 + 3 * e);
        const JT id = (JT)__ldcg(P.heavy + 3 + 3 * e);
        const int a = max(lo, slot_lo), b = min(lo + c, slot_hi);
        for (int sl = a + threadIdx.x; sl < b; sl += BLOCK) __stcg(jout_flat + (sl - P.first), id);
      }
    }
  }
  total_out = (num != Mslots) ? __ldcg(P.bins + P.n - 1) : 0.0;
  const int m0 = max(slot_lo, num);
  for (int m = m0 + threadIdx.x; m < slot_hi; m += BLOCK) {
    const unsigned k = (unsigned)(m - num);
    double u;
    if (u_draws) {
      u = __ldcg(u_draws + k);
    } else {
      const uint4 r = rng_block(P.key, ST_RESID, step_idx, (unsigned long long)k, 0);
      u = uniform53(r.x, r.y);
    }
    int lo = 0, hi = P.n;                                             // first i with u < bins[i]  (:108-113)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (u < __ldcg(P.bins + mid)) hi = mid;
      else lo = mid + 1;
    }
    if (lo < P.n) __stcg(jout_flat + (m - P.first), jbase + (JT)lo);
    else if (jid) __stcg(jout_flat + (m - P.first), jbase + (JT)m);
  }
  __syncthreads();
}

}  // namespace llpf
