// llpf_engine.cuh — the persistent, cooperative particle-filter engine for sm_100a.
//
// One launch runs a whole trajectory (or one step verb): the sequential time loop of
// forward_trajectory / loglik (reference src/filtering.jl:343-384, src/smoothing.jl:227-236) never
// returns to the host.  The grid is co-resident (cooperative launch); blocks own contiguous particle
// chunks and meet at a hand-rolled grid barrier (one red.release + ld.acquire spin per block).
//
// Pass structure for ParticleFilter / AdvancedParticleFilter (filtering.jl:140-168):
//     W(1)  [P(1)+W(2)] [P(2)+W(3)] ... [P(T-1)+W(T)]  P(T)
//   where W(k) = correct!(k)  (measurement_equation! PFtypes.jl:107-120 / :226-239 + logsumexp! utils.jl:18-27)
//         P(k) = predict!(k)  (shouldresample/resample resample.jl:5-36 + propagate_particles! PFtypes.jl:122-139,
//                              ext/...DistributionsExt.jl:83-93 + reset_weights! utils.jl:73-78)
//   predict!(k) and correct!(k+1) are fused into one sweep over the particles: 2*nx*8+16 B of traffic
//   per particle-step instead of the 3*nx*8+16 B of the un-fused formulation; the copyto!(xprev,x)
//   of filtering.jl:151 disappears (in-place update, ping-pong only on resample steps).
//
// Pass structure for AuxiliaryParticleFilter (filtering.jl:195-217): A(k) -> scan -> B(k), see aux_A/aux_B.
//
// Weight normalisation is lazy: w[] keeps the un-normalised log-weights and the pair (max, log sum)
// found by the grid reduction is applied when w is next read (same two subtractions as utils.jl:20,25).
// `we` is never materialised in the loop; ESS = S^2/Q comes out of the same reduction.
//
// Resampling (resample.jl:17-61): bins = device-wide inclusive scan of we.  FAST mode scans in 2^-62
// fixed point (u64): exact, associative, monotone, independent of block/GPU partitioning.  SERIAL mode
// is one thread doing the reference's left-to-right f64 adds (bit-exact verification mode).  Offspring
// indices come from a two-level upper-bound search (block table in smem, then the block's chunk).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "llpf_rng.cuh"

namespace llpf {

#ifndef LLPF_MIN_BLOCKS
#define LLPF_MIN_BLOCKS 3
#endif
constexpr int BLOCK = 256;
constexpr int NWARP = BLOCK / 32;
constexpr int MAX_BLOCKS = 1024;
constexpr int MAX_NU = 8;
constexpr int MAX_NX = 8;
constexpr int MAX_WORLD = 8;
constexpr int PS = 12;  // doubles per block partial: m, s, q, sx[8], pad
constexpr double FIX_SCALE = 4611686018427387904.0;        // 2^62
constexpr double FIX_INV = 2.168404344971008868e-19;       // 2^-62

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// persistent scalar state of one filter (device memory; the PFstate refs maxw/t plus bookkeeping)
// ------------------------------------------------------------------------------------------------
struct Scalars {
  long long t_index;      // index(pf) = state.t[]  (PFtypes.jl:16, filtering.jl:13,152)
  long long resample_count;
  int cur;                // which x buffer holds particles(pf)
  int uniform;            // 0: w[] materialised; 1: w == -log(N) (reset!, filtering.jl:11); 2: w == log(1/N) (reset_weights!, utils.jl:75)
  int pend;               // w[] un-normalised; logical w = (w - pend_m) - pend_ls
  int stats_ahead;        // APF: (pend_m,pend_ls,...) describe raw w[] but correct! has not been called yet
  int stats_valid;        // ess valid for the current weights
  int j_identity;         // state.j == 1:N (filtering.jl:148)
  int nonfinite;          // a log-likelihood increment was not finite
  int last_resampled;
  double pend_m, pend_ls, pend_s;
  double ess;
  double ll_last;
  double ll_total;
  double bins_total;
  double xhat[MAX_NX];
};

struct PeerTable {        // multi-GPU: peer views of each rank's arrays (IPC-mapped), index = rank
  double* x[MAX_WORLD][2];
  double* bins[MAX_WORLD];
  double* mailbox[MAX_WORLD];
};

struct EngineP {
  double* x[2];           // SoA: component d of local particle i at x[buf][d*ld + i]
  long long ld;
  double* w;              // [n] log-weights (lazy-normalised)
  double* lam;            // [n] APF lambda (the reference aliases state.we, filtering.jl:200)
  double* bins;           // [n] cumulative weights of the last resample (state.bins)
  int* j;                 // [n] 0-based global ancestor indices of the last resample (state.j)
  unsigned int* bar;      // grid barrier counter (zeroed by the host before every launch)
  double* partials;       // [MAX_BLOCKS*PS]
  u64* tots;              // [MAX_BLOCKS]
  Scalars* sc;
  const double* u;        // [T][nu]
  const double* y;        // [T][ny]
  double* ll_steps;       // optional per-step outputs (device)
  double* ess_steps;
  int* resampled;
  double* xhat;
  double* x_hist;         // [T][N][nx] AoS
  double* w_hist;         // [T][N]
  double* we_hist;        // [T][N]
  long long N;            // global particle count
  long long n;            // local particle count
  long long first;        // global index of local particle 0
  int T;                  // steps in this launch
  int prog;               // 0: PF program; 1: APF trajectory; 2: APF correct!; 3: APF predict!; 4: APF update!
  int lead_w;             // PF program starts with W(1)
  int lead_skip;          // ... which is reduce-only (no measurement update): refreshes ESS for predict!
  int trail_p;            // PF program ends with P(T)
  int filter;             // LLPF_FILTER_*
  int aux_tail_pf;        // APF loglik: last step is the inner filter's update! (smoothing.jl:235)
  int time_conv;          // 0: t=(k-1)*Ts (filtering.jl:352) ; 1: t=k*Ts (filtering.jl:181 with index from 1)
  int use_t_override;
  double t_override;
  double Ts;
  double thr;             // resample_threshold
  int strategy;           // LLPF_RESAMPLE_*
  int scan_mode;          // LLPF_SCAN_*
  int want_xhat;
  int nblocks;            // == gridDim.x
  long long chunk;        // particles per block
  RngKey key;
  int rank, world;
  double fix_scale, fix_inv;  // fixed-point scan scale: 2^62 / 2^-62 for normalised weights
};

template <int NX, int NY>
struct ModelP {
  double A[NX * NX];      // row-major
  double L1[NX * NX];     // row-major, lower Cholesky factor of R1 (noise = L1*z, utils.jl:260-268)
  double G[NY * NX];      // row-major: inv(chol(R2)) * C      (whitened measurement matrix)
  double W[NY * NY];      // row-major, lower: inv(chol(R2))
  double B[NX * MAX_NU];  // row-major NX x nu
  double c0;              // mvnormal_c0 = -(ny*log(2pi) + logdet R2)/2   (utils.jl:254-257)
  double qt[8];           // quadtank: {-a/A, -(a*f)/A, a/A, 2g, g1k1/A, g2k2/A, (1-g2)k2/A, (1-g1)k1/A}
  double t_switch;
  double integ_h;         // Ts0 / supersample  (utils.jl:223)
  int supersample;
  int nu;
};

// ------------------------------------------------------------------------------------------------
// synchronisation + block collectives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All blocks are co-resident (cooperative launch).  One arrive + spin per block.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    red_release_add_u32(bar, 1u);
    while ((int)(ld_acquire_u32(bar) - target) < 0) {
    }
    __threadfence();
  }
  __syncthreads();
}

struct Shared {
  double red[NWARP * (3 + MAX_NX)];
  u64 wtot[NWARP];
  double bu[MAX_NX];
  double yt[8];
  int skip;
  u64 offs[MAX_BLOCKS + 1];     // exclusive block offsets of the fixed-point scan
  double offd[MAX_BLOCKS + 1];  // the same as doubles == bins at chunk ends
};

// max over the block, result in every thread (deterministic)
__device__ __forceinline__ double block_max(double v, Shared& sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = sh.red[0];
#pragma unroll
  for (int i = 1; i < NWARP; ++i) r = fmax(r, sh.red[i]);
  return r;
}

// sums of M values over the block, results in every thread (fixed order => bitwise identical everywhere)
template <int M>
__device__ __forceinline__ void block_sum(double (&v)[M], Shared& sh) {
#pragma unroll
  for (int k = 0; k < M; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < M; ++k) sh.red[(threadIdx.x >> 5) * M + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < M; ++k) {
    double r = sh.red[k];
#pragma unroll
    for (int i = 1; i < NWARP; ++i) r += sh.red[i * M + k];
    v[k] = r;
  }
}

// online (max, sum exp, sum exp^2 [, sum exp*x]) accumulator: exactly one exp per sample
template <int NX>
struct Online {
  double m, s, q;
  double sx[NX];
  __device__ __forceinline__ void init() {
    m = -DBL_MAX; s = 0.0; q = 0.0;
#pragma unroll
    for (int d = 0; d < NX; ++d) sx[d] = 0.0;
  }
  __device__ __forceinline__ void add(double wv, const double (&x)[NX], bool with_x) {
    const double d = wv - m;
    const double e = exp(-fabs(d));
    if (d > 0.0) {   // new running maximum: rescale what we have
      s = fma(s, e, 1.0);
      q = fma(q, e * e, 1.0);
      if (with_x) {
#pragma unroll
        for (int k = 0; k < NX; ++k) sx[k] = fma(sx[k], e, x[k]);
      }
      m = wv;
    } else {
      s += e;
      q = fma(e, e, q);
      if (with_x) {
#pragma unroll
        for (int k = 0; k < NX; ++k) sx[k] = fma(e, x[k], sx[k]);
      }
    }
  }
};

struct Stats {
  double m, s, q;
  double sx[MAX_NX];
};

// Reduce the per-thread online accumulators to one block partial, publish it, meet the grid, and
// combine all block partials in a fixed order (every block computes bitwise-identical Stats).
template <int NX>
__device__ __forceinline__ Stats reduce_stats(const EngineP& P, Shared& sh, Online<NX>& acc,
                                              bool with_x, unsigned& bar_target) {
  const double mb = block_max(acc.m, sh);
  const double sc = exp(acc.m - mb);  // 0 for empty threads
  double v[2 + NX];
  v[0] = acc.s * sc;
  v[1] = acc.q * (sc * sc);
#pragma unroll
  for (int d = 0; d < NX; ++d) v[2 + d] = with_x ? acc.sx[d] * sc : 0.0;
  block_sum<2 + NX>(v, sh);
  if (threadIdx.x == 0) {
    double* p = P.partials + (size_t)blockIdx.x * PS;
    __stcg(p + 0, mb);
    __stcg(p + 1, v[0]);
    __stcg(p + 2, v[1]);
#pragma unroll
    for (int d = 0; d < NX; ++d) __stcg(p + 3 + d, v[2 + d]);
  }
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  // combine
  double m = -DBL_MAX;
  for (int b = threadIdx.x; b < P.nblocks; b += BLOCK) m = fmax(m, __ldcg(P.partials + (size_t)b * PS));
  m = block_max(m, sh);
  double t[2 + NX];
#pragma unroll
  for (int k = 0; k < 2 + NX; ++k) t[k] = 0.0;
  for (int b = threadIdx.x; b < P.nblocks; b += BLOCK) {
    const double* p = P.partials + (size_t)b * PS;
    const double e = exp(__ldcg(p) - m);
    t[0] = fma(__ldcg(p + 1), e, t[0]);
    t[1] = fma(__ldcg(p + 2), e * e, t[1]);
    if (with_x) {
#pragma unroll
      for (int d = 0; d < NX; ++d) t[2 + d] = fma(__ldcg(p + 3 + d), e, t[2 + d]);
    }
  }
  block_sum<2 + NX>(t, sh);
  Stats st;
  st.m = m; st.s = t[0]; st.q = t[1];
#pragma unroll
  for (int d = 0; d < NX; ++d) st.sx[d] = t[2 + d];
  return st;
}

// ------------------------------------------------------------------------------------------------
// scan + search (shared by the engine and the stand-alone resample entry points)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 to_fixed(double we, double scale) {
  // we*scale in [0,2^62] (normalised weights: scale = 2^62); NaN/negative -> 0
  const double v = we * scale;
  return (v > 0.0) ? __double2ull_rn(fmin(v, FIX_SCALE)) : 0ull;
}

// Stage 1: block-local inclusive scan of the block's chunk [beg,end).
// FAST  : bins[i] (as u64) = local inclusive fixed-point prefix; tots[b] = block total.
// SERIAL: bins[i] = we_i (double); the single-thread pass runs after the barrier.
template <class WeFn>
__device__ __forceinline__ void scan_stage1(const EngineP& P, Shared& sh, long long beg, long long end,
                                            WeFn wefn) {
  if (P.scan_mode != 0) {
    for (long long i = beg + threadIdx.x; i < end; i += BLOCK) __stcg(P.bins + i, wefn(i));
    return;
  }
  u64* lbins = reinterpret_cast<u64*>(P.bins);
  u64 carry = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = beg; base < end; base += BLOCK) {
    const long long i = base + threadIdx.x;
    u64 v = (i < end) ? to_fixed(wefn(i), P.fix_scale) : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) sh.wtot[warp] = v;
    __syncthreads();
    u64 woff = 0, rtot = 0;
#pragma unroll
    for (int k = 0; k < NWARP; ++k) {
      const u64 t = sh.wtot[k];
      if (k < warp) woff += t;
      rtot += t;
    }
    if (i < end) __stcg(lbins + i, carry + woff + v);
    carry += rtot;
  }
  if (threadIdx.x == 0) __stcg(P.tots + blockIdx.x, carry);
}

// Stage 2 (after a grid barrier): exclusive block offsets, then finalise bins for the own chunk.
// Leaves sh.offs / sh.offd filled for FAST mode.  `base_fixed` is this GPU's global CDF offset.
__device__ __forceinline__ void scan_stage2(const EngineP& P, Shared& sh, long long beg, long long end,
                                            u64 base_fixed) {
  const int nb = P.nblocks;
  // every block scans all block totals (exact integer arithmetic: any order gives the same bits)
  __syncthreads();
  u64 t4[4];
  u64 mine = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int b = threadIdx.x * 4 + k;
    t4[k] = (b < nb) ? __ldcg(P.tots + b) : 0ull;
    mine += t4[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 v = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u64 t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) sh.wtot[warp] = v;
  __syncthreads();
  u64 woff = 0;
#pragma unroll
  for (int k = 0; k < NWARP; ++k)
    if (k < warp) woff += sh.wtot[k];
  u64 excl = base_fixed + woff + v - mine;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int b = threadIdx.x * 4 + k;
    if (b <= nb) {
      sh.offs[b] = excl;
      sh.offd[b] = (double)excl * P.fix_inv;
    }
    excl += t4[k];
  }
  __syncthreads();
  const u64 off = sh.offs[blockIdx.x];
  u64* lbins = reinterpret_cast<u64*>(P.bins);
  for (long long i = beg + threadIdx.x; i < end; i += BLOCK) {
    const u64 loc = __ldcg(lbins + i);
    __stcg(P.bins + i, (double)(off + loc) * P.fix_inv);
  }
}

// SERIAL mode: the reference's cumsum (resample.jl:19-22), one thread, strict left-to-right f64 adds.
__device__ __noinline__ void scan_serial(double* bins, long long n) {
  double acc = __ldcg(bins);
  long long i = 1;
  for (; i + 8 <= n; i += 8) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(bins + i + k);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc = __dadd_rn(acc, v[k]);
      __stcg(bins + i + k, acc);
    }
  }
  for (; i < n; ++i) {
    acc = __dadd_rn(acc, __ldcg(bins + i));
    __stcg(bins + i, acc);
  }
}

// After the bins barrier in SERIAL mode: block table = bins at chunk ends.
__device__ __forceinline__ void load_block_table(const EngineP& P, Shared& sh) {
  __syncthreads();
  for (int b = threadIdx.x; b <= P.nblocks; b += BLOCK) {
    if (b == 0) {
      sh.offd[0] = 0.0;
    } else {
      long long e = (long long)b * P.chunk;
      if (e > P.n) e = P.n;
      sh.offd[b] = __ldcg(P.bins + e - 1);
    }
  }
  __syncthreads();
}

// smallest local index a with bins[a] > s  (== the reference's two-pointer search, resample.jl:26-34,
// expressed as an upper bound).  Requires s < offd[nb].  Two levels: block table, then the chunk.
__device__ __forceinline__ long long upper_bound_bins(const EngineP& P, const Shared& sh, double s) {
  int lo = 0, hi = P.nblocks - 1;  // find smallest b with offd[b+1] > s
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sh.offd[mid + 1] > s) hi = mid; else lo = mid + 1;
  }
  long long a = (long long)lo * P.chunk;
  long long e = a + P.chunk;
  if (e > P.n) e = P.n;
  long long b = e - 1;  // bins[b] == offd[lo+1] > s
  while (a < b) {
    const long long mid = (a + b) >> 1;
    if (__ldcg(P.bins + mid) > s) b = mid; else a = mid + 1;
  }
  return a;
}

struct Thresholds {
  double r, step, total, M;
  int strategy;
  const double* u_slots;   // stratified: caller-supplied rand() per slot (stand-alone entry), else RNG
};

// s[i] = fl(r + fl(i * fl(1/M)))   (Julia StepRangeLen getindex; resample.jl:24) — the
// product and the sum are rounded separately (no FMA).  Stratified: ((i + rand_i)/M)*bins[N] (resample.jl:49).
__device__ __forceinline__ double threshold(const Thresholds& th, const RngKey& key, uint32_t step_idx,
                                            long long gi) {
  if (th.strategy == 1) {
    double u;
    if (th.u_slots) {
      u = th.u_slots[gi];
    } else {
      const uint4 r = rng_block(key, ST_STRAT, step_idx, (unsigned long long)gi, 0);
      u = uniform53(r.x, r.y);
    }
    return __dmul_rn(__ddiv_rn(__dadd_rn((double)gi, u), th.M), th.total);
  }
  return __dadd_rn(th.r, __dmul_rn((double)gi, th.step));
}

__device__ __forceinline__ Thresholds make_thresholds(const EngineP& P, double total, double u01,
                                                      double Mslots, const double* u_slots) {
  Thresholds th;
  th.total = total;
  th.M = Mslots;
  th.strategy = P.strategy;
  th.u_slots = u_slots;
  th.step = __ddiv_rn(1.0, Mslots);
  // r = rand()*bins[end]/N   (resample.jl:23; note /N, N = length(we))
  th.r = __ddiv_rn(__dmul_rn(u01, total), (double)P.N);
  return th;
}
__device__ __forceinline__ double resample_u01(const RngKey& key, uint32_t step_idx) {
  const uint4 r = rng_block(key, ST_RESAMPLE, step_idx, 0ull, 0);
  return uniform53(r.x, r.y);
}

// ------------------------------------------------------------------------------------------------
// models
// ------------------------------------------------------------------------------------------------
// quadtank right-hand side, example_quadtank.jl:91-106 (+ the t>500 leak switch of :15-17)
template <int NX, int NY>
__device__ __forceinline__ void quadtank_rhs(const ModelP<NX, NY>& M, const double* bu, double t,
                                             const double (&h)[NX], double (&xd)[NX]) {
  const double c1 = (t > M.t_switch) ? M.qt[1] : M.qt[0];
  const double co = M.qt[0], ci = M.qt[2], tg = M.qt[3];
  const double s0 = sqrt(fmax(tg * h[0], 0.0) + 1e-3);
  const double s1 = sqrt(fmax(tg * h[1], 0.0) + 1e-3);
  const double s2 = sqrt(fmax(tg * h[2], 0.0) + 1e-3);
  const double s3 = sqrt(fmax(tg * h[3], 0.0) + 1e-3);
  xd[0] = c1 * s0 + ci * s2 + bu[0];
  xd[1] = co * s1 + ci * s3 + bu[1];
  xd[2] = co * s2 + bu[2];
  xd[3] = co * s3 + bu[3];
}

// dynamics(x,u,p,t) without noise, in place.  DYN==0: A*x .+ B*u ; DYN==1: rk4(quadtank) utils.jl:220-237
template <int NX, int NY, int DYN>
__device__ __forceinline__ void dynamics_mean(const ModelP<NX, NY>& M, const double* bu, double t,
                                              double (&x)[NX]) {
  if (DYN == 0) {
    double xn[NX];
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      double acc = M.A[r * NX] * x[0];
#pragma unroll
      for (int c = 1; c < NX; ++c) acc = fma(M.A[r * NX + c], x[c], acc);
      xn[r] = acc + bu[r];
    }
#pragma unroll
    for (int r = 0; r < NX; ++r) x[r] = xn[r];
  } else {
    const double h = M.integ_h;
    for (int s = 0; s < M.supersample; ++s) {
      double f1[NX], f2[NX], f3[NX], f4[NX], tmp[NX];
      quadtank_rhs<NX, NY>(M, bu, t, x, f1);
#pragma unroll
      for (int i = 0; i < NX; ++i) tmp[i] = x[i] + h / 2 * f1[i];
      quadtank_rhs<NX, NY>(M, bu, t + h / 2, tmp, f2);
#pragma unroll
      for (int i = 0; i < NX; ++i) tmp[i] = x[i] + h / 2 * f2[i];
      quadtank_rhs<NX, NY>(M, bu, t + h / 2, tmp, f3);
#pragma unroll
      for (int i = 0; i < NX; ++i) tmp[i] = x[i] + h * f3[i];
      quadtank_rhs<NX, NY>(M, bu, t + h, tmp, f4);
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] += h / 6 * (f1[i] + 2 * f2[i] + 2 * f3[i] + f4[i]);
      t += h;
    }
  }
}

// x += L1*z, z ~ N(0,I) from the (ST_DYN, step, particle) counter  (PFtypes.jl:135,153; utils.jl:260-268)
template <int NX, int NY>
__device__ __forceinline__ void add_dynamics_noise(const ModelP<NX, NY>& M, const RngKey& key,
                                                   uint32_t step_idx, long long gi, double (&x)[NX]) {
  double z[NX];
  normals<NX>(key, ST_DYN, step_idx, (unsigned long long)gi, z);
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c <= r; ++c) acc = fma(M.L1[r * NX + c], z[c], acc);
    x[r] += acc;
  }
}

// logpdf(N(0,R2), y - C x) = c0 - |W(y - Cx)|^2/2 = c0 - |yt - G x|^2/2   (utils.jl:252-257)
template <int NX, int NY>
__device__ __forceinline__ double meas_loglik(const ModelP<NX, NY>& M, const double* yt,
                                              const double (&x)[NX]) {
  double q = 0.0;
#pragma unroll
  for (int a = 0; a < NY; ++a) {
    double v = yt[a];
#pragma unroll
    for (int c = 0; c < NX; ++c) v = fma(-M.G[a * NX + c], x[c], v);
    q = fma(v, v, q);
  }
  return fma(-0.5, q, M.c0);
}

// ------------------------------------------------------------------------------------------------
// the engine
// ------------------------------------------------------------------------------------------------
template <int NX, int NY, int DYN>
struct Engine {
  const EngineP& P;
  const ModelP<NX, NY>& M;
  Shared& sh;
  Scalars sc;          // every block carries an identical copy, block 0 writes it back at the end
  unsigned bar_target;
  long long beg, end;  // own chunk (local indices)
  double lwN;          // -log(N)  (filtering.jl:11)
  double lw1N;         // log(1/N) (utils.jl:75)

  __device__ Engine(const EngineP& p, const ModelP<NX, NY>& m, Shared& s) : P(p), M(m), sh(s) {
    sc = *P.sc;
    bar_target = 0;
    beg = (long long)blockIdx.x * P.chunk;
    end = beg + P.chunk;
    if (end > P.n) end = P.n;
    if (beg > P.n) beg = P.n;
    lwN = -log((double)P.N);
    lw1N = log(1.0 / (double)P.N);
  }

  __device__ __forceinline__ double step_time(int k) const {  // k is 1-based
    if (P.use_t_override) return P.t_override;
    return (double)(k - 1 + P.time_conv) * P.Ts;
  }

  // per-pass uniform data: bu = B*u_k (or the quadtank input terms), yt = W*y_k, skip = any(isnan(y))
  __device__ __forceinline__ void stage_step(int k_u, int k_y) {
    __syncthreads();
    if (k_u > 0 && threadIdx.x < NX) {
      const double* u = P.u + (size_t)(k_u - 1) * M.nu;
      double acc = 0.0;
      if (DYN == 0) {
        for (int c = 0; c < M.nu; ++c) acc = fma(M.B[threadIdx.x * MAX_NU + c], u[c], acc);
      } else {
        // {g1k1/A*u1, g2k2/A*u2, (1-g2)k2/A*u2, (1-g1)k1/A*u1}
        const int ui = (threadIdx.x == 0 || threadIdx.x == 3) ? 0 : 1;
        acc = M.qt[4 + threadIdx.x] * u[ui];
      }
      sh.bu[threadIdx.x] = acc;
    }
    if (k_y > 0 && threadIdx.x == 32) {
      const double* y = P.y + (size_t)(k_y - 1) * NY;
      int skip = 0;
#pragma unroll
      for (int a = 0; a < NY; ++a) {
        if (isnan(y[a])) skip = 1;
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c <= a; ++c) acc = fma(M.W[a * NY + c], y[c], acc);
        sh.yt[a] = acc;
      }
      sh.skip = skip;
    }
    __syncthreads();
  }

  __device__ __forceinline__ void load_x(const double* buf, long long i, double (&x)[NX]) const {
#pragma unroll
    for (int d = 0; d < NX; ++d) x[d] = buf[(size_t)d * P.ld + i];
  }
  __device__ __forceinline__ void load_x_cg(const double* buf, long long i, double (&x)[NX]) const {
#pragma unroll
    for (int d = 0; d < NX; ++d) x[d] = __ldcg(buf + (size_t)d * P.ld + i);
  }
  __device__ __forceinline__ void store_x(double* buf, long long i, const double (&x)[NX]) const {
#pragma unroll
    for (int d = 0; d < NX; ++d) buf[(size_t)d * P.ld + i] = x[d];
  }
  __device__ __forceinline__ void store_hist_x(int k, long long gi, const double (&x)[NX]) const {
    double* p = P.x_hist + ((size_t)(k - 1) * P.N + gi) * NX;
#pragma unroll
    for (int d = 0; d < NX; ++d) __stcs(p + d, x[d]);
  }

  // by-value snapshot of the lazy weight state (keeps the particle loops free of this->sc reloads)
  struct WState {
    int uniform, pend;
    double pm, pls, inv_s, wu, weu;
    const double* w;
    // logical (normalised) log-weight of local particle i: (w - offset) - log1p(s)  utils.jl:20,25
    __device__ __forceinline__ double weight_norm(long long i) const {
      if (uniform) return wu;
      const double wr = w[i];
      return pend ? (wr - pm) - pls : wr;
    }
    // we = exp(w - offset) * 1/(s+1)   utils.jl:21-24
    __device__ __forceinline__ double expweight(long long i) const {
      if (uniform) return weu;
      const double wr = w[i];
      return pend ? exp(wr - pm) * inv_s : exp(wr);
    }
  };
  __device__ __forceinline__ WState wstate() const {
    WState ws;
    ws.uniform = sc.uniform; ws.pend = sc.pend;
    ws.pm = sc.pend_m; ws.pls = sc.pend_ls;
    ws.inv_s = sc.pend ? 1.0 / sc.pend_s : 1.0;
    ws.wu = (sc.uniform == 1) ? lwN : lw1N;
    ws.weu = 1.0 / (double)P.N;
    ws.w = P.w;
    return ws;
  }

  // ---- resampling: bins <- scan(we); returns the thresholds --------------------------------------
  template <class WeFn>
  __device__ __forceinline__ Thresholds build_bins(WeFn wefn) {
    scan_stage1(P, sh, beg, end, wefn);
    grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
    if (P.scan_mode != 0) {
      if (blockIdx.x == 0 && threadIdx.x == 0) scan_serial(P.bins, P.n);
      grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
      load_block_table(P, sh);
    } else {
      scan_stage2(P, sh, beg, end, 0ull);
      grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
    }
    const double total = sh.offd[P.nblocks];
    sc.bins_total = total;
    return make_thresholds(P, total, resample_u01(P.key, (uint32_t)sc.t_index), (double)P.N, nullptr);
  }

  // ancestor (local==global index, world==1) of output slot i; stale slots keep state.j (resample.jl:26-34)
  __device__ __forceinline__ long long ancestor(const Thresholds& th, long long i) const {
    const double s = threshold(th, P.key, (uint32_t)sc.t_index, P.first + i);
    if (s < th.total) return upper_bound_bins(P, sh, s);
    return sc.j_identity ? i : (long long)P.j[i];
  }

  __device__ __forceinline__ void publish_step(int k, const Stats& st) {
    const double ls = log(st.s);
    const double ll = st.m + ls;  // log1p(s)+offset of utils.jl:26 (s there excludes the arg-max term)
    sc.pend = 1; sc.uniform = 0; sc.stats_ahead = 0; sc.stats_valid = 1;
    sc.pend_m = st.m; sc.pend_ls = ls; sc.pend_s = st.s;
    sc.ess = st.s * st.s / st.q;   // effective_particles = 1/sum(abs2, we)  resample.jl:1-2
    sc.ll_last = ll;
    sc.ll_total += ll;
    if (!(fabs(ll) <= DBL_MAX)) sc.nonfinite = 1;
#pragma unroll
    for (int d = 0; d < NX; ++d) sc.xhat[d] = st.sx[d] / st.s;
    if (blockIdx.x == 0 && threadIdx.x == 0 && k > 0) {
      if (P.ll_steps) P.ll_steps[k - 1] = ll;
      if (P.ess_steps) P.ess_steps[k - 1] = sc.ess;
      if (P.xhat) {
#pragma unroll
        for (int d = 0; d < NX; ++d) P.xhat[(size_t)(k - 1) * NX + d] = sc.xhat[d];
      }
    }
  }

  // shouldresample(pf)  resample.jl:5-10
  __device__ __forceinline__ bool should_resample() const {
    if (P.thr == 1.0) return true;
    return sc.ess < (double)P.N * P.thr;
  }

  // ---- PF / AdvancedPF pass: [predict!(k_prop)] fused with [correct!(k_weigh)] ---------------------
  // k_prop / k_weigh are 1-based step numbers, 0 = phase absent.  skip_meas: reduce-only correct!.
  __device__ __noinline__ void pf_pass(int k_prop, int k_weigh, bool skip_meas) {
    stage_step(k_prop, skip_meas ? 0 : k_weigh);
    const bool skip = skip_meas || (k_weigh > 0 && sh.skip);
    const bool res = (k_prop > 0) && should_resample();
    const bool hist_w = (P.w_hist != nullptr) && (k_prop > 0);   // weights of step k_prop, just corrected
    const WState ws = wstate();
    const EngineP& Pr = P;
    Thresholds th;
    if (res) {
      th = build_bins([=, &Pr](long long i) {
        const double we = ws.expweight(i);
        if (hist_w) {
          const size_t o = (size_t)(k_prop - 1) * Pr.N + Pr.first + i;
          __stcs(Pr.w_hist + o, ws.weight_norm(i));
          __stcs(Pr.we_hist + o, we);
        }
        return we;
      });
    }
    const double* src = P.x[sc.cur];
    double* dst = res ? P.x[sc.cur ^ 1] : P.x[sc.cur];
    const double tprop = step_time(k_prop);
    const bool with_x = (P.want_xhat != 0);
    Online<NX> acc;
    acc.init();
    for (long long i = beg + threadIdx.x; i < end; i += BLOCK) {
      const long long gi = P.first + i;
      double x[NX];
      double wv;
      if (res) {
        const long long a = ancestor(th, i);
        load_x_cg(src, a, x);
        P.j[i] = (int)a;
        wv = lw1N;  // reset_weights!  utils.jl:75
      } else {
        load_x(src, i, x);
        wv = ws.weight_norm(i);
        if (hist_w) {
          const size_t o = (size_t)(k_prop - 1) * P.N + gi;
          __stcs(P.w_hist + o, wv);
          __stcs(P.we_hist + o, ws.expweight(i));
        }
      }
      if (k_prop > 0) {
        dynamics_mean<NX, NY, DYN>(M, sh.bu, tprop, x);
        add_dynamics_noise<NX, NY>(M, P.key, (uint32_t)sc.t_index, gi, x);
        store_x(dst, i, x);
      }
      if (k_weigh > 0) {
        if (P.x_hist) store_hist_x(k_weigh, gi, x);
        if (!skip) wv += meas_loglik<NX, NY>(M, sh.yt, x);
        P.w[i] = wv;
        acc.add(wv, x, with_x);
      }
    }
    if (k_prop > 0) {
      if (res) {
        sc.cur ^= 1;
        sc.uniform = 2; sc.pend = 0; sc.stats_ahead = 0;
        sc.ess = (double)P.N; sc.stats_valid = 1;
        sc.j_identity = 0;
        sc.resample_count += 1;
      } else {
        sc.j_identity = 1;   // s.j .= 1:N  filtering.jl:148
      }
      sc.last_resampled = res ? 1 : 0;
      if (blockIdx.x == 0 && threadIdx.x == 0 && P.resampled) P.resampled[k_prop - 1] = res ? 1 : 0;
      sc.t_index += 1;       // filtering.jl:152
    }
    if (k_weigh > 0) {
      const Stats st = reduce_stats<NX>(P, sh, acc, with_x, bar_target);
      publish_step(k_weigh, st);
    }
  }

  // ---- AuxiliaryParticleFilter predict!(pfa,u,y1,p,t)  filtering.jl:195-217 (and :219-234) ----------
  // A: x̄ = f(x) (no noise) ; λ = logpdf(y1 - C x̄) ; v = w + λ ; expnormalize!(v)  -> W
  // scan(W) -> bins -> j
  // B: x = x̄[j] + L z ; w = λ - log N (UNresampled λ, :210-213) ; stats of the new w == next correct!
  __device__ __noinline__ void aux_step(int k, int k_y1) {
    stage_step(k, k_y1);
    const bool skip = sh.skip;
    const bool adv = (P.filter == 3);
    const double tprop = step_time(k);
    const double* cur = P.x[sc.cur];
    double* oth = P.x[sc.cur ^ 1];
    const bool hist_w = (P.w_hist != nullptr);
    const WState ws = wstate();
    Online<NX> acc;
    acc.init();
    for (long long i = beg + threadIdx.x; i < end; i += BLOCK) {
      double x[NX];
      load_x(cur, i, x);
      const double wn = ws.weight_norm(i);
      if (hist_w) {
        const size_t o = (size_t)(k - 1) * P.N + P.first + i;
        __stcs(P.w_hist + o, wn);
        __stcs(P.we_hist + o, ws.expweight(i));
      }
      dynamics_mean<NX, NY, DYN>(M, sh.bu, tprop, x);           // :199 no noise
      if (!adv) store_x(oth, i, x);
      const double lam = skip ? 0.0 : meas_loglik<NX, NY>(M, sh.yt, x);  // :200-202
      P.lam[i] = lam;
      const double v = wn + lam;                                 // :203
      P.w[i] = v;
      acc.add(v, x, false);
    }
    const Stats s1 = reduce_stats<NX>(P, sh, acc, false, bar_target);
    const double inv1 = 1.0 / s1.s;
    // expnormalize!(w): exp(w-offset)*1/(s+1)   utils.jl:57-63 ; then resample (always)  :205
    const double m1 = s1.m;
    const double* wraw = P.w;
    const Thresholds th = build_bins([=](long long i) { return exp(wraw[i] - m1) * inv1; });
    const bool with_x = (P.want_xhat != 0);
    const double lN = log((double)P.N);
    acc.init();
    for (long long i = beg + threadIdx.x; i < end; i += BLOCK) {
      const long long gi = P.first + i;
      const long long a = ancestor(th, i);
      P.j[i] = (int)a;
      double x[NX];
      double wnew;
      if (adv) {
        load_x_cg(cur, a, x);                                    // :230 propagate again from xprev[j]
        dynamics_mean<NX, NY, DYN>(M, sh.bu, tprop, x);
        add_dynamics_noise<NX, NY>(M, P.key, (uint32_t)sc.t_index, gi, x);
        store_x(oth, i, x);
        wnew = lw1N;                                             // :228 reset_weights!
      } else {
        load_x_cg(oth, a, x);                                    // :207 permute_with_buffer!
        add_dynamics_noise<NX, NY>(M, P.key, (uint32_t)sc.t_index, gi, x);  // :208 add_noise!
        store_x(P.x[sc.cur], i, x);
        wnew = P.lam[i] - lN;                                    // :210-213
      }
      if (P.x_hist) store_hist_x(k + 1, gi, x);
      P.w[i] = wnew;
      acc.add(wnew, x, with_x);
    }
    if (adv) sc.cur ^= 1;
    sc.j_identity = 0;
    sc.resample_count += 1;
    sc.last_resampled = 1;
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.resampled) P.resampled[k - 1] = 1;
    sc.t_index += 1;                                             // :215
    const Stats s2 = reduce_stats<NX>(P, sh, acc, with_x, bar_target);
    // stats of the raw w[] are ready; correct! (filtering.jl:170-174) has not been *called* yet
    sc.pend = 0; sc.uniform = 0;
    sc.stats_ahead = 1; sc.stats_valid = 0;
    sc.pend_m = s2.m; sc.pend_ls = log(s2.s); sc.pend_s = s2.s;
    sc.ess = s2.s * s2.s / s2.q;
#pragma unroll
    for (int d = 0; d < NX; ++d) sc.xhat[d] = s2.sx[d] / s2.s;
  }

  // correct!(pfa,...) = logsumexp!(state) only  filtering.jl:170-174, using the stats found by aux_step
  __device__ __forceinline__ void aux_correct_from_stats(int k) {
    const double ll = sc.pend_m + sc.pend_ls;
    sc.pend = 1; sc.stats_ahead = 0; sc.stats_valid = 1;
    sc.ll_last = ll;
    sc.ll_total += ll;
    if (!(fabs(ll) <= DBL_MAX)) sc.nonfinite = 1;
    if (blockIdx.x == 0 && threadIdx.x == 0 && k > 0) {
      if (P.ll_steps) P.ll_steps[k - 1] = ll;
      if (P.ess_steps) P.ess_steps[k - 1] = sc.ess;
      if (P.xhat) {
#pragma unroll
        for (int d = 0; d < NX; ++d) P.xhat[(size_t)(k - 1) * NX + d] = sc.xhat[d];
      }
    }
  }

  // write the weight history of step k for filters whose last step has no following pass
  __device__ void flush_weight_history(int k) {
    if (!P.w_hist) return;
    const WState ws = wstate();
    for (long long i = beg + threadIdx.x; i < end; i += BLOCK) {
      const size_t o = (size_t)(k - 1) * P.N + P.first + i;
      __stcs(P.w_hist + o, ws.weight_norm(i));
      __stcs(P.we_hist + o, ws.expweight(i));
    }
  }

  // correct!(pfa): reuse the stats found by aux_step when they describe the current weights,
  // otherwise a reduce-only sweep (logsumexp! of whatever the weights are)
  __device__ __forceinline__ void aux_correct(int k) {
    if (sc.stats_ahead) aux_correct_from_stats(k);
    else pf_pass(0, k, true);
  }

  __device__ void run() {
    const int T = P.T;
    switch (P.prog) {
      case 0: {  // ParticleFilter / AdvancedParticleFilter:  W(1) [P(k)+W(k+1)]... P(T)
        if (P.lead_w) pf_pass(0, 1, P.lead_skip != 0);
        for (int k = 1; k < T; ++k) pf_pass(k, k + 1, false);
        if (P.trail_p) pf_pass(T, 0, false);
      } break;
      case 1: {  // forward_trajectory(pfa) filtering.jl:367-384 / loglik(pfa) smoothing.jl:232-236
        const bool tail = (P.aux_tail_pf != 0);
        if (!(tail && T == 1)) aux_correct(1);
        for (int k = 1; k < T; ++k) {
          aux_step(k, k + 1);
          if (k + 1 < T || !tail) aux_correct_from_stats(k + 1);
        }
        if (tail) {
          // pf.pf(u[end], y[end], p, (T-1)*Ts): the INNER filter's update! — its correct! adds the
          // likelihood of y[T] on top of the raw (un-normalised) w = λ - log N, then predict!.
          if (sc.stats_ahead) { sc.pend = 0; sc.stats_ahead = 0; }
          pf_pass(0, T, false);
          pf_pass(T, 0, false);
        } else {
          flush_weight_history(T);
        }
      } break;
      case 2: aux_correct(1); break;                       // correct!(pfa)
      case 3: aux_step(1, 1); break;                       // predict!(pfa,u,y1): y1 staged at y[0]
      case 4: aux_correct(1); aux_step(1, 2); break;       // update!(pfa,u,y,y1): y1 staged at y[1]
      default: break;
    }
    // every block must have read the incoming scalars before block 0 overwrites them
    grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
    if (blockIdx.x == 0 && threadIdx.x == 0) *P.sc = sc;
  }
};

template <int NX, int NY, int DYN>
__global__ void __launch_bounds__(BLOCK, LLPF_MIN_BLOCKS)
k_engine(const __grid_constant__ EngineP P, const __grid_constant__ ModelP<NX, NY> M) {
  __shared__ Shared sh;
  Engine<NX, NY, DYN> e(P, M, sh);
  e.run();
}

}  // namespace llpf
