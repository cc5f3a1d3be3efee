// llpf_smooth.cuh — particle smoother, forward filtering / backward simulation (FFBS)
// reference: smooth(pf, xf, wf, wef, ll, M, u, y, p)  src/smoothing.jl:116-143, draw_one_categorical src/resample.jl:128-152.
//
//   xb[m,T] = xf[j[m],T],  j = resample(strategy, wef[:,T], M)
//   for t = T-1 ... 1, for every trajectory m:
//       wb[n] = wf[n,t] + logpdf(df, xb[m,t+1] - f(xf[n,t], u[t], p, (t-1)Ts))       n = 1..N
//       i     = first b with  rand() * sum_n exp(wb[n])  <=  sum_{n<=b} exp(wb[n])     (draw_one_categorical)
//       xb[m,t] = xf[i,t]
// O(M N T) transition-density evaluations; the M trajectories are independent given the stored forward pass.
//
// Device formulation: one 512-thread block per trajectory m (ordinary launch, no grid synchronisation), looping over
// the time steps backwards.  Per (m, t):
//   level 1  every warp owns a contiguous 1/16 of the particles: one streaming pass (coalesced AoS rows, dynamics,
//            whitened residual, online max / sum-exp with exactly one exp per particle) -> 16 (max, sum) pairs;
//   level 2  the warp range that contains the drawn quantile is re-evaluated by all 16 warps with the now known
//            global maximum -> 16 exact sub-sums;
//   level 3  one warp walks the selected sub-range (N/256 particles) with 32-wide shuffle scans to the index.
// ~1.07 N evaluations per draw instead of the reference's 3 passes + serial cumsum + linear search; the forward history
// slab of step t (N rows) is shared by all M blocks through L2.
// Equal to the reference's draw except when the quantile falls within rounding distance of a bin edge.
#pragma once
#include "llpf_engine.cuh"

namespace llpf {

constexpr int SM_BLOCK = 512;
constexpr int SM_WARPS = SM_BLOCK / 32;

struct SmoothP {
  const double* xf;       // [T][N][nx]  filtered particles  (ParticleFilteringSolution.x, AoS)
  const double* wf;       // [T][N]      normalised log-weights (sol.w)
  const double* u;        // [T][nu]
  const long long* j0;    // [M] 1-based indices of the resample at T (0 = slot left untouched)
  double* xb;             // [T][M][nx]  smoothed trajectories (M x T Matrix{SVector}, column-major)
  int N, M, T;
  double Ts;
  double c0;              // mvnormal_c0 of the dynamics density: -(nx log 2pi + logdet R1)/2   utils.jl:254-257
  RngKey key;
};

// model in dimension-independent form (row-major, stride MAX_NX); llpf_api.cu fills it
struct SmoothModelG {
  double A[MAX_NX * MAX_NX];
  double Winv[MAX_NX * MAX_NX];   // inverse of the lower Cholesky factor of R1 (whitening: invquad = |Winv r|^2)
  double B[MAX_NX * MAX_NU];
  double qt[8];
  double t_switch, integ_h;
  int supersample, nu;
};

struct SmoothShared {
  double xb[MAX_NX];
  double wm[SM_WARPS], ws[SM_WARPS];
  int pick;
};

template <int NX, int DYN>
__global__ void __launch_bounds__(SM_BLOCK)
k_smooth(const __grid_constant__ SmoothP S, const __grid_constant__ ModelP<NX, 1> Mo) {
  __shared__ Shared sh;
  __shared__ SmoothShared ss;
  math_tab_load(sh.mt);
  model_to_shared<NX, 1>(Mo, sh);   // sh.mA = A, sh.mL = Winv
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = S.N;
  // warp ranges: multiples of 32 particles
  const int per1 = (((N + SM_WARPS - 1) / SM_WARPS) + 31) & ~31;
  for (int m = blockIdx.x; m < S.M; m += gridDim.x) {
    if (threadIdx.x < NX) {
      long long a = S.j0[m];
      if (a < 1 || a > N) a = N;
      const double v = __ldg(S.xf + ((size_t)(S.T - 1) * N + (size_t)(a - 1)) * NX + threadIdx.x);
      ss.xb[threadIdx.x] = v;
      S.xb[((size_t)(S.T - 1) * S.M + m) * NX + threadIdx.x] = v;
    }
    for (int t = S.T - 1; t >= 1; --t) {   // the reference's 1-based t   smoothing.jl:130
      __syncthreads();                     // ss.xb of step t+1 is in place; previous step's scratch is free
      if (threadIdx.x < NX) {              // B u_t (or the quadtank input terms), as in stage_step
        const double* u = S.u + (size_t)(t - 1) * Mo.nu;
        double acc = 0.0;
        if (DYN == 0) {
          for (int c = 0; c < Mo.nu; ++c) acc = fma(Mo.B[threadIdx.x * MAX_NU + c], __ldg(u + c), acc);
        } else {
          const int ui = (threadIdx.x == 0 || threadIdx.x == 3) ? 0 : 1;
          acc = Mo.qt[4 + threadIdx.x] * __ldg(u + ui);
        }
        sh.bu[threadIdx.x] = acc;
      }
      __syncthreads();
      double bu[NX], xbn[NX];
#pragma unroll
      for (int r = 0; r < NX; ++r) { bu[r] = sh.bu[r]; xbn[r] = ss.xb[r]; }
      const double ti = (double)(t - 1) * S.Ts;   // :131
      const double* xft = S.xf + (size_t)(t - 1) * N * NX;
      const double* wft = S.wf + (size_t)(t - 1) * N;
      // wb[n] = wf[n,t] + logpdf(df, xb[m,t+1] - f(xf[n,t],u[t],p,ti))   :135 — split into the loads and the arithmetic so
      // that the sweeps can put two particles' loads in flight before either is consumed (the kernel stalled 55 % of
      // its time on these loads when each row was fetched right before use: profiles/r1_v9_ncu_smooth.md)
      struct Row { double x[NX]; double w; };
      auto load_row = [&](int n) -> Row {
        Row r;
        const double* row = xft + (size_t)n * NX;
        if constexpr (NX % 2 == 0) {
#pragma unroll
          for (int d = 0; d < NX; d += 2) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(row + d));
            r.x[d] = v.x; r.x[d + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int d = 0; d < NX; ++d) r.x[d] = __ldg(row + d);
        }
        r.w = __ldg(wft + n);
        return r;
      };
      auto wb_eval = [&](Row r) -> double {
        dynamics_mean<NX, 1, DYN>(Mo, sh, bu, ti, r.x);
        double q = 0.0;
#pragma unroll
        for (int rr = 0; rr < NX; ++rr) {
          double l[NX];
          lds_row<NX>(sh.mL + rr * mdl_stride(NX), l);
          double acc = l[0] * (xbn[0] - r.x[0]);
#pragma unroll
          for (int c = 1; c <= rr; ++c) acc = fma(l[c], xbn[c] - r.x[c], acc);
          q = fma(acc, acc, q);
        }
        return r.w + fma(-0.5, q, S.c0);
      };
      auto wb_of = [&](int n) -> double { return wb_eval(load_row(n)); };
      // ---- level 1: per-warp (max, sum exp) over contiguous ranges ------------------------------------------
      const int b1 = min(N, warp * per1), e1 = min(N, b1 + per1);
      Online<1> acc;
      acc.init();
      const double dummy[1] = {0.0};
      for (int n = b1 + lane; n < e1; n += 64) {   // two particles per iteration: both rows in flight, two chains
        const bool two = (n + 32 < e1);
        const Row ra = load_row(n);
        const Row rb = load_row(two ? n + 32 : n);
        const double wa = wb_eval(ra), wbv = wb_eval(rb);
        acc.add(wa, dummy, false, sh.mt);
        if (two) acc.add(wbv, dummy, false, sh.mt);
      }
      double wm = acc.m;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wm = fmax(wm, __shfl_xor_sync(0xffffffffu, wm, o));
      double wsum = acc.s * exp_nonpos(acc.m - wm, sh.mt);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
      if (lane == 0) { ss.wm[warp] = wm; ss.ws[warp] = wsum; }
      __syncthreads();
      double gm = ss.wm[0];
#pragma unroll
      for (int k = 1; k < SM_WARPS; ++k) gm = fmax(gm, ss.wm[k]);
      double tot = 0.0;
      double pre1[SM_WARPS];
#pragma unroll
      for (int k = 0; k < SM_WARPS; ++k) {
        tot += ss.ws[k] * exp_nonpos(ss.wm[k] - gm, sh.mt);
        pre1[k] = tot;
      }
      const uint4 rr = rng_block(S.key, ST_SMOOTH, (uint32_t)t, (unsigned long long)(unsigned)m, 0);
      const double target = uniform53(rr.x, rr.y) * tot;   // s = rand()*bins[end]   resample.jl:135
      int w1 = SM_WARPS - 1;
#pragma unroll
      for (int k = SM_WARPS - 1; k >= 0; --k)
        if (target <= pre1[k]) w1 = k;
      while (w1 > 0 && min(N, w1 * per1) >= N) --w1;       // never an empty range
      double off = 0.0;
#pragma unroll
      for (int k = 0; k < SM_WARPS; ++k)
        if (k == w1 - 1) off = pre1[k];
      // ---- level 2: the selected range, split over all warps, exact sums relative to the global maximum -----
      const int rb = min(N, w1 * per1), re = min(N, rb + per1);
      const int per2 = ((((re - rb) + SM_WARPS - 1) / SM_WARPS) + 31) & ~31;
      const int b2 = min(re, rb + warp * per2), e2 = min(re, b2 + per2);
      double s2 = 0.0;
      for (int n = b2 + lane; n < e2; n += 64) {
        const bool two = (n + 32 < e2);
        const Row ra = load_row(n);
        const Row rb = load_row(two ? n + 32 : n);
        const double ea = exp_nonpos(wb_eval(ra) - gm, sh.mt), eb = exp_nonpos(wb_eval(rb) - gm, sh.mt);
        s2 += ea;
        if (two) s2 += eb;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      __syncthreads();                                     // everyone has consumed ss.ws of level 1
      if (lane == 0) ss.ws[warp] = s2;
      __syncthreads();
      if (warp == 0) {
        // ---- level 3: warp 0 walks the selected sub-range -----------------------------------------------------
        int w2 = SM_WARPS - 1;
        double run = off, off2 = off;
        bool found = false;
#pragma unroll
        for (int k = 0; k < SM_WARPS; ++k) {
          const double nxt = run + ss.ws[k];
          if (!found && target <= nxt) { w2 = k; off2 = run; found = true; }
          run = nxt;
        }
        if (!found) {                                      // rounding: the re-evaluated range came up short
          w2 = SM_WARPS - 1;
          while (w2 > 0 && min(re, rb + w2 * per2) >= re) --w2;
          off2 = off;
          for (int k = 0; k < w2; ++k) off2 += ss.ws[k];
        }
        const int b3 = min(re, rb + w2 * per2), e3 = min(re, b3 + per2);
        int pick = e3 - 1;
        double carry = off2;
        for (int base = b3; base < e3; base += 32) {
          const int n = base + lane;
          double e = (n < e3) ? exp_nonpos(wb_of(n) - gm, sh.mt) : 0.0;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, e, o);
            if (lane >= o) e += v;
          }
          const unsigned hit = __ballot_sync(0xffffffffu, (n < e3) && (target <= carry + e));
          if (hit) { pick = base + (__ffs(hit) - 1); break; }
          carry += __shfl_sync(0xffffffffu, e, 31);
        }
        if (lane == 0) ss.pick = pick;
      }
      __syncthreads();
      if (threadIdx.x < NX) {                              // xb[m,t] = xf[i,t]   :139
        const double v = __ldg(xft + (size_t)ss.pick * NX + threadIdx.x);
        S.xb[((size_t)(t - 1) * S.M + m) * NX + threadIdx.x] = v;
        ss.xb[threadIdx.x] = v;
      }
    }
    __syncthreads();
  }
}

// host entry of the translation unit llpf_smooth.cu
cudaError_t smooth_launch(int nx, int dyn, const SmoothP& S, const SmoothModelG& G, int grid, cudaStream_t stream);

}  // namespace llpf
