// llpf_stats.cuh — weighted statistics of a forward history that stays in HBM:
//   mean_trajectory / mode_trajectory   src/filtering.jl:417-440   (sum(x .* we), particle with the largest weight)
//   weighted_cov                        src/filtering.jl:575-583   (StatsBase cov, ProbabilityWeights, corrected = true)
//   weighted_quantile                   src/filtering.jl:592-595   (StatsBase quantile with ProbabilityWeights)
// The reference computes these on the host from the N x T matrices of the ParticleFilteringSolution; at config 2 that
// history is ~50 GB, so it is reduced where it lives and only T x (...) results cross PCIe.
// History layout (llpf_engine.cuh): x [T][N][nx] AoS, we [T][N].  One thread block per time step (moments) or per
// (time step, component, quantile) (quantiles); fixed summation order => run-to-run deterministic.
#pragma once
#include "llpf_engine.cuh"

namespace llpf {

// mean, mode, covariance of one time step per block.  cov: two passes (mean first, then centred second moments) like
// StatsBase; factor n / ((n - 1) sum(w)) with n = number of non-zero weights (ProbabilityWeights, corrected = true).
template <int NX>
__global__ void __launch_bounds__(BLOCK)
k_hist_moments(const double* __restrict__ xh, const double* __restrict__ weh, long long N, int T,
               double* __restrict__ xmean, double* __restrict__ xmode, double* __restrict__ xcov) {
  constexpr int RS = NX * NX + NX + 4;   // row stride of the per-warp partials
  __shared__ double red[NWARP * RS];
  __shared__ double bc[NX + 2];
  __shared__ long long bidx;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const double* x = xh + (size_t)t * N * NX;
    const double* we = weh + (size_t)t * N;
    // ---- pass 1: sum(we), sum(we x), count(we != 0), first arg-max of we ----
    double s = 0.0, sx[NX], cnt = 0.0, best = -1.0;
    long long bi = 0x7fffffffffffffffll;
#pragma unroll
    for (int d = 0; d < NX; ++d) sx[d] = 0.0;
    for (long long n = threadIdx.x; n < N; n += BLOCK) {
      const double e = __ldcs(we + n);
      s += e;
      cnt += (e != 0.0) ? 1.0 : 0.0;
      if (e > best) { best = e; bi = n; }          // strided ascending n per thread: the first maximum of the thread
#pragma unroll
      for (int d = 0; d < NX; ++d) sx[d] = fma(e, __ldcs(x + (size_t)n * NX + d), sx[d]);
    }
    // block reduction (fixed order)
    double v[NX + 2];
    v[0] = s; v[1] = cnt;
#pragma unroll
    for (int d = 0; d < NX; ++d) v[2 + d] = sx[d];
#pragma unroll
    for (int k = 0; k < NX + 2; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // findmax: largest weight, smallest index among equals (Julia findmax)
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      for (int k = 0; k < NX + 2; ++k) red[(threadIdx.x >> 5) * RS + k] = v[k];
      red[(threadIdx.x >> 5) * RS + NX + 2] = best;
      red[(threadIdx.x >> 5) * RS + NX + 3] = __longlong_as_double(bi);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot[NX + 2];
      for (int k = 0; k < NX + 2; ++k) tot[k] = 0.0;
      double b = -1.0;
      long long i0 = 0x7fffffffffffffffll;
      for (int w = 0; w < NWARP; ++w) {
        for (int k = 0; k < NX + 2; ++k) tot[k] += red[w * RS + k];
        const double ob = red[w * RS + NX + 2];
        const long long oi = __double_as_longlong(red[w * RS + NX + 3]);
        if (ob > b || (ob == b && oi < i0)) { b = ob; i0 = oi; }
      }
      bc[0] = tot[0]; bc[1] = tot[1];
      for (int d = 0; d < NX; ++d) bc[2 + d] = tot[2 + d] / tot[0];     // weighted mean of the step
      bidx = i0;
    }
    __syncthreads();
    const double sw = bc[0], nn = bc[1];
    double mu[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) mu[d] = bc[2 + d];
    if (threadIdx.x < NX) {
      if (xmean) xmean[(size_t)t * NX + threadIdx.x] = mu[threadIdx.x] * sw;   // mean_trajectory: sum(x .* we), unnormalised
      if (xmode) xmode[(size_t)t * NX + threadIdx.x] = x[(size_t)bidx * NX + threadIdx.x];
    }
    if (xcov) {
      // ---- pass 2: centred second moments ----
      double c[NX * NX];
#pragma unroll
      for (int k = 0; k < NX * NX; ++k) c[k] = 0.0;
      for (long long n = threadIdx.x; n < N; n += BLOCK) {
        const double e = __ldcs(we + n);
        double dx[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) dx[d] = __ldcs(x + (size_t)n * NX + d) - mu[d];
#pragma unroll
        for (int a = 0; a < NX; ++a)
#pragma unroll
          for (int b2 = 0; b2 <= a; ++b2) c[a * NX + b2] = fma(e * dx[a], dx[b2], c[a * NX + b2]);
      }
#pragma unroll
      for (int a = 0; a < NX; ++a)
#pragma unroll
        for (int b2 = 0; b2 <= a; ++b2)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) c[a * NX + b2] += __shfl_xor_sync(0xffffffffu, c[a * NX + b2], o);
      __syncthreads();
      if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < NX * NX; ++k) red[(threadIdx.x >> 5) * RS + k] = c[k];
      __syncthreads();
      if (threadIdx.x < NX * NX) {
        const int a = threadIdx.x / NX, b2 = threadIdx.x % NX;
        const int k = (b2 <= a) ? a * NX + b2 : b2 * NX + a;
        double tot = 0.0;
        for (int w = 0; w < NWARP; ++w) tot += red[w * RS + k];
        xcov[(size_t)t * NX * NX + threadIdx.x] = tot * (nn / ((nn - 1.0) * sw));
      }
    }
    __syncthreads();
  }
}

// order-preserving map double -> u64
__device__ __forceinline__ u64 f64_key(double v) {
  const u64 b = (u64)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(u64 k) {
  const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// StatsBase.quantile(v, ProbabilityWeights(w), p) for component `comp` of time step t: one block per (t, comp, q).
//   drop w == 0 ; sort by v ; h = p (sum(w) - w_1) + w_1 ; walk the cumulative weights S_k while S_k <= h ;
//   result = v_{k-1} + (h - S_{k-1}) / (S_k - S_{k-1}) (v_k - v_{k-1})                      (k > N: v_N)
// No sort: the crossing element is found by a most-significant-digit-first radix SELECT on the order-preserving key of
// v, 8 rounds of 8 bits, each a weighted 256-bin histogram of the elements that still match the prefix.  Weights are
// accumulated in 2^-62 fixed point (exact, order-independent sums), so the result does not depend on thread scheduling.
__global__ void __launch_bounds__(BLOCK)
k_hist_quantile(const double* __restrict__ xh, const double* __restrict__ weh, long long N, int T, int nx,
                const double* __restrict__ qs, int nq, double* __restrict__ out /*[T][nq][nx]*/) {
  __shared__ u64 hist[256];
  __shared__ u64 s_pref, s_below, s_kmin, s_w1, s_tot, s_wk, s_kprev, s_kmax;
  const long long items = (long long)T * nx * nq;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int qi = (int)(item % nq);
    const int comp = (int)((item / nq) % nx);
    const int t = (int)(item / ((long long)nq * nx));
    const double* x = xh + (size_t)t * N * nx + comp;
    const double* we = weh + (size_t)t * N;
    // ---- pass 0: total weight, smallest / largest key among the non-zero weights, weight of the smallest ----
    u64 tot = 0, kmin = ~0ull, kmax = 0ull;
    for (long long n = threadIdx.x; n < N; n += BLOCK) {
      const double e = __ldcs(we + n);
      if (e != 0.0) {   // StatsBase drops exactly-zero weights only: a weight of 1e-300 still takes part in the order
        const u64 k = f64_key(__ldcs(x + (size_t)n * nx));
        tot += to_fixed(e, FIX_SCALE);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot += __shfl_xor_sync(0xffffffffu, tot, o);
      const u64 a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
      kmin = a < kmin ? a : kmin;
      kmax = b > kmax ? b : kmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_tot = 0; s_kmin = ~0ull; s_kmax = 0ull; s_w1 = 0; s_wk = 0; s_kprev = 0; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      atomicAdd((unsigned long long*)&s_tot, (unsigned long long)tot);
      atomicMin((unsigned long long*)&s_kmin, (unsigned long long)kmin);
      atomicMax((unsigned long long*)&s_kmax, (unsigned long long)kmax);
    }
    __syncthreads();
    const u64 KMIN = s_kmin, KMAX = s_kmax, TOT = s_tot;
    {
      u64 w1 = 0;
      for (long long n = threadIdx.x; n < N; n += BLOCK) {
        const double e = __ldcs(we + n);
        if (e != 0.0 && f64_key(__ldcs(x + (size_t)n * nx)) == KMIN) w1 += to_fixed(e, FIX_SCALE);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w1 += __shfl_xor_sync(0xffffffffu, w1, o);
      if ((threadIdx.x & 31) == 0 && w1) atomicAdd((unsigned long long*)&s_w1, (unsigned long long)w1);
    }
    __syncthreads();
    const double p = qs[qi];
    const double W = (double)TOT * FIX_INV, w1d = (double)s_w1 * FIX_INV;
    const double h = p * (W - w1d) + w1d;
    const double hs = h * FIX_SCALE;
    const u64 ht = (hs >= 9.2e18) ? ~0ull : __double2ull_rn(hs);
    double result;
    if (KMIN > KMAX) {
      result = __longlong_as_double(0x7ff8000000000000ll);       // all weights zero: NaN
    } else if (p == 0.0 || TOT == 0) {
      result = key_f64(KMIN);                                    // h = w_1: v_1 + 0 * (v_2 - v_1)
    } else if (p == 1.0 || ht >= TOT) {
      result = key_f64(KMAX);                                    // S_k <= h for every k: v[end]
    } else {
      // ---- radix select: smallest key K with S(<= K) > ht ----
      if (threadIdx.x == 0) { s_pref = 0; s_below = 0; }
      __syncthreads();
      for (int round = 0; round < 8; ++round) {
        const int shift = 56 - 8 * round;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        const u64 pref = s_pref;
        for (long long n = threadIdx.x; n < N; n += BLOCK) {
          const u64 wf = to_fixed(__ldcs(we + n), FIX_SCALE);
          if (!wf) continue;
          const u64 k = f64_key(__ldcs(x + (size_t)n * nx));
          if (round == 0 || (k >> (shift + 8)) == pref) atomicAdd((unsigned long long*)&hist[(k >> shift) & 255], (unsigned long long)wf);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          u64 acc = s_below;
          int d = 0;
          for (; d < 255; ++d) {
            if (acc + hist[d] > ht) break;
            acc += hist[d];
          }
          s_below = acc;
          s_pref = (pref << 8) | (u64)d;
          if (round == 7) s_wk = hist[d];
        }
        __syncthreads();
      }
      const u64 K = s_pref;
      // ---- the element before it in sorted order: largest key < K among the non-zero weights ----
      u64 kp = 0;
      bool any = false;
      for (long long n = threadIdx.x; n < N; n += BLOCK) {
        if (__ldcs(we + n) == 0.0) continue;
        const u64 k = f64_key(__ldcs(x + (size_t)n * nx));
        if (k < K && (!any || k > kp)) { kp = k; any = true; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const u64 ok = __shfl_xor_sync(0xffffffffu, kp, o);
        const bool oa = __shfl_xor_sync(0xffffffffu, (int)any, o) != 0;
        if (oa && (!any || ok > kp)) { kp = ok; any = true; }
      }
      if ((threadIdx.x & 31) == 0 && any) atomicMax((unsigned long long*)&s_kprev, (unsigned long long)kp);
      __syncthreads();
      const double vk = key_f64(K);
      const double Skold = (double)s_below * FIX_INV, Sk = (double)(s_below + s_wk) * FIX_INV;
      if (s_kprev == 0 && !(KMIN < K)) {
        result = vk;                                             // the crossing element is the first one
      } else {
        const double vkold = key_f64(s_kprev);
        result = vkold + (h - Skold) / (Sk - Skold) * (vk - vkold);
      }
    }
    if (threadIdx.x == 0) out[((size_t)t * nq + qi) * nx + comp] = result;
    __syncthreads();
  }
}

}  // namespace llpf
