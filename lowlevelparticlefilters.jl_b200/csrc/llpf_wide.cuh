// llpf_wide.cuh — the engine for WIDE linear-Gaussian models with Float32 particles (BASELINE config 5:
// ParticleFilter, 64-state LG in the regime of test/test_large.jl:8-22, N=2^20, Float32 particles, Float64
// weights — the reference keeps w/we/bins in Float64 whatever the particle eltype, src/PFtypes.jl:68-69).
//
// Same persistent cooperative design, op list, lazy weight normalisation, reduction and resampling as
// llpf_engine.cuh (those device functions are reused verbatim); what differs is the per-particle work:
//   propagate   x' = A x + B u + L z     64x64 mat-vec per particle   (PFtypes.jl:122-139, ext/...DistributionsExt.jl:83-93)
//   weigh       w += c0 - |yt - G x'|^2/2,   G = chol(R2)^-1 C  (ny x 64)  (PFtypes.jl:107-120, utils.jl:252-257)
// i.e. ~8k FP32 FMAs against 784 B per particle-step: FP32-FMA-bound, not HBM-bound.
//
// Layout: particles are AoS in HBM, 64 consecutive floats (256 B = two full lines) per particle — the reference's
// Vector{SVector} layout — so the ancestor gather of a resampling step reads whole lines.  One thread owns one
// particle: its 64 accumulators live in registers as 32 packed f32x2 pairs and are updated with FFMA2
// (fma.rn.f32x2, sm_100+), the matrix columns are broadcast from shared memory with 16-byte loads
// (A^T, L^T column-major; G row-major).  Noise: the same Philox4x32-10 / f64 Box-Muller contract as the f64
// engine, 16 counter blocks per particle, rounded to f32.
//
// Summation order (ours: the reference would call BLAS sgemv, whose order is unspecified) — identical in
// oracle/llpf_oracle.c (f32 section), every operation a correctly rounded fmaf / addition in f32:
//   acc_r = 0 ; for c = 0..63: acc_r = fmaf(A[r,c], x[c], acc_r) ; acc_r += (B u)_r ; for c <= r: acc_r = fmaf(L[r,c], z[c], acc_r)
//   d_a   = (sum over even c, fmaf chain from 0) + (sum over odd c, fmaf chain from 0) ; v = yt_a - d_a
//   q     = fmaf(v, v, q) for a ascending ; loglik = fmaf(-0.5, q, c0)  (f32) ; w += (double)loglik
#pragma once
#include "llpf_wide_common.cuh"

namespace llpf {

// dynamic shared memory of the wide kernel (v3: block-tiled, 128-thread blocks, two per SM)
constexpr int WTP = BLOCK;     // particles per tile (one per thread in the load / store / reduce phases)
static_assert(WTP == 128, "the thread <-> tile mapping below assumes 128-thread blocks (llpf_wide.cu)");
struct WideShared {
  alignas(16) float As[WNX * WNX];    // column-major A:  As[c*64 + r] = A[r,c]   (two consecutive rows = one f32x2 operand)
  alignas(16) float Gs[WNX * WNX];    // column-major G:  Gs[c*64 + a] = G[a,c]   (rows a >= ny are zero)
  alignas(16) float Ls[WNX * WNX];    // column-major lower Cholesky factor of R1 (zero above the diagonal)
  alignas(16) float XT[WNX * WTP];    // tile buffer, component-major: XT[c*WTP + p]; x, then z (general L), x', v in turn
  alignas(16) float bu[WNX];
  alignas(16) float yt[WNX];
  alignas(16) float ldiag[WNX];
};

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo);
}
__device__ __forceinline__ float lo_f(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi_f(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void lds_2x64(const float* p, u64& a, u64& b) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
// 32-byte global accesses (one full sector per thread), L2-coherent
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float (&v)[8]) {
  asm volatile("st.global.cg.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(p)
               : "memory");
}

// where particle `a` (GLOBAL index) lives: buffer buf_id of the owning rank (peer memory when sharded)
__device__ __forceinline__ const float* wide_row(const EngineP& P, int buf_id, int a) {
  if (P.world > 1) {
    const int r = a / P.n;
    return reinterpret_cast<const float*>(P.peer_x[r][buf_id]) + (size_t)(a - r * P.n) * WNX;
  }
  return reinterpret_cast<const float*>(P.x[buf_id]) + (size_t)(a - P.first) * WNX;
}

// per-pass uniform data: bu = B u_k (f32 fmaf chain), yt = (float)(W y_k), skip = any(isnan(y))  (PFtypes.jl:109)
__device__ __forceinline__ void wide_stage_step(const EngineP& P, const WideP& Mw, Shared& sh, WideShared& ws, int k_u,
                                                int k_y, bool& skip) {
  __syncthreads();
  if (k_u > 0 && threadIdx.x < WNX) {
    const double* u = P.u + (size_t)(k_u - 1) * Mw.nu;
    float acc = 0.f;
    for (int c = 0; c < Mw.nu; ++c) acc = fmaf(Mw.B[threadIdx.x * MAX_NU + c], (float)__ldg(u + c), acc);
    ws.bu[threadIdx.x] = acc;
  }
  if (k_y > 0 && threadIdx.x >= WNX && threadIdx.x < 2 * WNX) {
    const int a = threadIdx.x - WNX;
    const double* y = P.y + (size_t)(k_y - 1) * Mw.ny;
    double acc = 0.0;
    if (a < Mw.ny)
      for (int c = 0; c <= a; ++c) acc = fma(Mw.W[a * Mw.ny + c], __ldg(y + c), acc);
    ws.yt[a] = (float)acc;
    if (a == 0) {
      int sk = 0;
      for (int c = 0; c < Mw.ny; ++c)
        if (isnan(__ldg(y + c))) sk = 1;
      sh.skip = sk;
    }
  }
  __syncthreads();
  skip = (k_y > 0) ? (sh.skip != 0) : false;
}

// ---- PF pass for wide models: [predict!(k_prop)] fused with [correct!(k_weigh)]  (cf. pf_pass) ------------
// v1 gave every thread one particle and its 64 accumulators and streamed all three matrices from shared memory for every
// particle: 16 LDS.128 per 32 FFMA2 — shared-memory bound at 27-29 % of the FP32 peak (profiles/r1_v9_ncu_wide.md).
// v3 (v2 was the same with one 256-thread block per SM) processes a tile of 128 particles per block iteration as two
// 64 x 64 x 128 GEMMs on FFMA2:
//   thread t = 16 rg + pg owns rows 8rg..8rg+7 of particles {4pg..4pg+3} and {64+4pg..64+4pg+3}: 64 accumulators as 32
//   f32x2 pairs (two consecutive ROWS per pair); per column c it loads 8 matrix elements (2 LDS.128, two addresses per
//   warp) and 8 particle values (2 conflict-free LDS.128) for 32 FFMA2: 4 LDS.128 instead of 16 per 32 FFMA2.
// Every accumulator still sees exactly the fmaf chain documented at the top of this file (c ascending from 0, then + Bu,
// then the noise chain; even / odd column chains for G x'), so the results are bit-identical to v1 and to the oracle.
// Tile phases (one __syncthreads between them): gather x -> XT | GEMM1 A x (+Bu, +L z) , store x' | x' -> XT | GEMM2 G x',
// v = yt - d | v -> XT | per particle: q = sum v^2 (fmaf chain over a), w += c0 - q/2, online reduction.
// Two such blocks are resident per SM and run out of step, so one block's FP64 Box-Muller / global-load latency overlaps
// the other's FFMA2 phases.

// 4 consecutive floats of MUTABLE shared data (bu, yt: rewritten every pass) as two f32x2 pairs.  lds_2x64 is a
// non-volatile asm without memory clobber — fine for the matrices, which never change after the prologue, but the
// compiler may move such a load across the __syncthreads that orders it after the writer.
__device__ __forceinline__ void lds_pairs_mutable(const float* p, u64& a, u64& b) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  a = pack2(v.x, v.y);
  b = pack2(v.z, v.w);
}

// acc[rp][p] += M[8rg + 2rp .. +1, c] * t[c][particle p]   for c = c0, c0 + cs, ... < 64
// p: 0..3 = particles 4pg..4pg+3, 4..7 = particles 64+4pg..64+4pg+3 (HALF selects 0..3, 4..7 or all eight)
template <int HALF>
__device__ __forceinline__ void tile_gemm(u64 (&acc)[4][8], const float* Mc, const float* XT, int rg, int pg, int c0, int cs) {
#pragma unroll 2
  for (int c = c0; c < WNX; c += cs) {
    u64 m[4];
    lds_2x64(Mc + c * WNX + 8 * rg, m[0], m[1]);
    lds_2x64(Mc + c * WNX + 8 * rg + 4, m[2], m[3]);
    float x[8];
    if (HALF != 1) {
      const float4 v = *reinterpret_cast<const float4*>(XT + c * WTP + 4 * pg);
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
    if (HALF != 0) {
      const float4 v = *reinterpret_cast<const float4*>(XT + c * WTP + 64 + 4 * pg);
      x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w;
    }
#pragma unroll
    for (int p = (HALF == 1 ? 4 : 0); p < (HALF == 0 ? 4 : 8); ++p) {
      const u64 xx = pack2(x[p], x[p]);
#pragma unroll
      for (int rp = 0; rp < 4; ++rp) acc[rp][p] = ffma2(m[rp], xx, acc[rp][p]);
    }
  }
}

__device__ __forceinline__ void pf_pass_wide(const EngineP& P, const WideP& Mw, Shared& sh, WideShared& ws,
                                             Scalars& sc, Ctx& cx, int k_prop, int k_weigh, int flags) {
  if (flags & OPF_RAW_WEIGHTS) { sc.pend = 0; sc.stats_ahead = 0; }
  const bool skip_meas = (flags & OPF_SKIP_MEAS) != 0;
  bool nan_y;
  wide_stage_step(P, Mw, sh, ws, k_prop, skip_meas ? 0 : k_weigh, nan_y);
  const bool skip = skip_meas || nan_y;
  const bool res = (k_prop > 0) && ((P.thr == 1.0) || (sc.ess < (double)P.N * P.thr));   // resample.jl:5-10
  const WState wst = make_wstate(P, sc, cx, sh.mt);
  const uint32_t step_idx = (uint32_t)sc.t_index;
  int f_total = 0;
  if (res && P.strategy == 2) {   // ResampleResidual  resample.jl:63-117
    WeSrc rs;
    rs.w = wst.w; rs.mode = wst.uniform ? 1 : (wst.pend ? 3 : 2);
    rs.pm = wst.pm; rs.pls = wst.pls; rs.inv_s = wst.inv_s; rs.weu = wst.weu; rs.wu = wst.wu; rs.T = &sh.mt;
    rs.hist_w = nullptr; rs.hist_we = nullptr;
    double total;
    resample_residual<int>(P, sh, cx.beg, cx.end, cx.bar_target, rs, nullptr, step_idx, (int)P.N, P.j, P.first,
                           sc.j_identity, P.first + cx.beg, P.first + cx.end, total);
    f_total = (int)P.N;
    sc.bins_total = total;
  } else if (res) {
    double total;
    f_total = resample_indices<int>(
        P, sh, cx.beg, cx.end, cx.bar_target,
        [=](int i) { return wst.uniform ? 0.0 : __ldcg(wst.w + i); },
        [=](int, double wr) { return wst.expweight_raw(wr); },
        0.0, true, step_idx, (int)P.N, nullptr, P.j, P.first, total, sc.xseq, P.first + cx.beg, P.first + cx.end,
        nullptr, sc.cur);
    sc.bins_total = total;
  }
  const int cur = sc.cur;
  float* dst = reinterpret_cast<float*>(res ? P.x[cur ^ 1] : P.x[cur]);
  const int jid = sc.j_identity;
  const int rg = threadIdx.x >> 4, pg = threadIdx.x & 15;
  const bool prop = k_prop > 0, weigh = k_weigh > 0;
  Online<1> acc1;
  acc1.init();
  const double dummy[1] = {0.0};
  for (int tb = cx.beg; tb < cx.end; tb += WTP) {   // block-uniform trip count
    // ---- phase 0: this thread's particle of the tile: ancestor, old weight, row -> XT[:, p] ----
    const int i = tb + threadIdx.x;
    const bool valid = i < cx.end;
    const int gi = P.first + i;
    double wraw = 0.0;
    __syncthreads();                               // the previous tile's readers of XT are done
    {
      float xv[8];
      const float* xin = nullptr;
      if (valid && (prop || weigh)) {
        int a = gi;
        if (res) {   // ancestor (resample.jl:26-34: slots past the last threshold keep state.j)
          a = __ldcg(P.j + i);
          if (gi >= f_total) {
            if (jid) a = gi;
            __stcg(P.j + i, a);
          }
        }
        if (!res && !wst.uniform) wraw = __ldcg(P.w + i);
        if (P.world > 1 && a < 0) {   // an ancestor another rank shipped here as a packed entry (expand_packs)
          const char* e = P.pack_in + (size_t)(-1 - a) * (size_t)P.pack_stride;
          xin = reinterpret_cast<const float*>(e);
          __stcg(P.j + i, __ldcg(reinterpret_cast<const int*>(e + P.pack_state_bytes)));   // state.j keeps global ids
        } else {
          xin = wide_row(P, cur, prop ? a : gi);
        }
      }
#pragma unroll
      for (int c8 = 0; c8 < WNX / 8; ++c8) {
        if (xin) ldg256(xin + 8 * c8, xv);
        else {
#pragma unroll
          for (int k = 0; k < 8; ++k) xv[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) ws.XT[(8 * c8 + k) * WTP + threadIdx.x] = xv[k];
      }
    }
    __syncthreads();
    // tile-local particle index of accumulator column p
    auto pcol = [&](int p) { return ((p < 4) ? 0 : 64) + 4 * pg + (p & 3); };
    if (prop) {
      // ---- phase 1: x' = A x (+ B u) (+ L z), rows 8rg..8rg+7 of 8 particles ----
      u64 acc[4][8];
#pragma unroll
      for (int rp = 0; rp < 4; ++rp)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[rp][p] = 0ull;
      tile_gemm<2>(acc, ws.As, ws.XT, rg, pg, 0, 1);
      {   // + B u
        u64 b[4];
        lds_pairs_mutable(ws.bu + 8 * rg, b[0], b[1]);
        lds_pairs_mutable(ws.bu + 8 * rg + 4, b[2], b[3]);
#pragma unroll
        for (int rp = 0; rp < 4; ++rp)
#pragma unroll
          for (int p = 0; p < 8; ++p) acc[rp][p] = fadd2(acc[rp][p], b[rp]);
      }
      // noise: 2 Philox blocks (rows 8rg..8rg+3, 8rg+4..8rg+7) per particle, f64 Box-Muller rounded to f32 — the same
      // counters as v1: block b of particle gi holds the normals of rows 4b..4b+3
      if (Mw.diagL) {
        u64 ld[4];
        lds_2x64(ws.ldiag + 8 * rg, ld[0], ld[1]);
        lds_2x64(ws.ldiag + 8 * rg + 4, ld[2], ld[3]);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int pgi = P.first + tb + pcol(p);
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const uint4 rr = rng_block(P.key, ST_DYN, step_idx, (unsigned long long)(unsigned)pgi, (uint32_t)(2 * rg + bb));
            const uint32_t ra[2] = {rr.x, rr.z}, rb[2] = {rr.y, rr.w};
            double a0[2], a1[2];
            normal_pairs<2>(ra, rb, a0, a1, sh.mt);
            // rows 8rg + 4bb + {0,1} and {2,3}
            acc[2 * bb][p] = ffma2(ld[2 * bb], pack2((float)a0[0], (float)a1[0]), acc[2 * bb][p]);
            acc[2 * bb + 1][p] = ffma2(ld[2 * bb + 1], pack2((float)a0[1], (float)a1[1]), acc[2 * bb + 1][p]);
          }
        }
      } else {
        // general L: z tile through XT (thread p generates all 64 normals of its particle), then the chain
        // acc_r = fmaf(L[r,c], z[c], acc_r) over c ascending (L is zero above the diagonal: those terms add +0 exactly)
        __syncthreads();                           // GEMM1 readers of XT are done
#pragma unroll 1
        for (int b = 0; b < WNX / 4; ++b) {
          const uint4 rr = rng_block(P.key, ST_DYN, step_idx, (unsigned long long)(unsigned)gi, (uint32_t)b);
          const uint32_t ra[2] = {rr.x, rr.z}, rb[2] = {rr.y, rr.w};
          double a0[2], a1[2];
          normal_pairs<2>(ra, rb, a0, a1, sh.mt);
          ws.XT[(4 * b) * WTP + threadIdx.x] = (float)a0[0];
          ws.XT[(4 * b + 1) * WTP + threadIdx.x] = (float)a1[0];
          ws.XT[(4 * b + 2) * WTP + threadIdx.x] = (float)a0[1];
          ws.XT[(4 * b + 3) * WTP + threadIdx.x] = (float)a1[1];
        }
        __syncthreads();
        tile_gemm<2>(acc, ws.Ls, ws.XT, rg, pg, 0, 1);
      }
      // store x' (rows 8rg..8rg+7 = one 32-byte sector per particle) and hand the tile to phase 2 through XT
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int ip = tb + pcol(p);
        if (ip < cx.end) {
          float o[8];
#pragma unroll
          for (int rp = 0; rp < 4; ++rp) { o[2 * rp] = lo_f(acc[rp][p]); o[2 * rp + 1] = hi_f(acc[rp][p]); }
          stg256(dst + (size_t)ip * WNX + 8 * rg, o);
        }
      }
      if (weigh) {
        __syncthreads();                           // readers of XT (x or z) are done
#pragma unroll
        for (int rp = 0; rp < 4; ++rp) {
          float* r0 = ws.XT + (8 * rg + 2 * rp) * WTP;
          float* r1 = r0 + WTP;
          *reinterpret_cast<float4*>(r0 + 4 * pg) = make_float4(lo_f(acc[rp][0]), lo_f(acc[rp][1]), lo_f(acc[rp][2]), lo_f(acc[rp][3]));
          *reinterpret_cast<float4*>(r1 + 4 * pg) = make_float4(hi_f(acc[rp][0]), hi_f(acc[rp][1]), hi_f(acc[rp][2]), hi_f(acc[rp][3]));
          *reinterpret_cast<float4*>(r0 + 64 + 4 * pg) = make_float4(lo_f(acc[rp][4]), lo_f(acc[rp][5]), lo_f(acc[rp][6]), lo_f(acc[rp][7]));
          *reinterpret_cast<float4*>(r1 + 64 + 4 * pg) = make_float4(hi_f(acc[rp][4]), hi_f(acc[rp][5]), hi_f(acc[rp][6]), hi_f(acc[rp][7]));
        }
        __syncthreads();
      }
    }
    double wv;
    if (res) wv = cx.lw1N;                    // reset_weights!  utils.jl:75
    else wv = wst.uniform ? wst.wu : (wst.pend ? (wraw - wst.pm) - wst.pls : wraw);
    if (weigh) {
      if (!skip) {
        // ---- phase 2: d_a = (even-column chain) + (odd-column chain) of G[a,:] x' ; v_a = yt_a - d_a ----
        float v[8][8];   // [row 8rg + r][accumulator column p]
        u64 yp[4];
        lds_pairs_mutable(ws.yt + 8 * rg, yp[0], yp[1]);
        lds_pairs_mutable(ws.yt + 8 * rg + 4, yp[2], yp[3]);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          u64 de[4][8], dd[4][8];
#pragma unroll
          for (int rp = 0; rp < 4; ++rp)
#pragma unroll
            for (int p = 0; p < 8; ++p) { de[rp][p] = 0ull; dd[rp][p] = 0ull; }
          if (half == 0) { tile_gemm<0>(de, ws.Gs, ws.XT, rg, pg, 0, 2); tile_gemm<0>(dd, ws.Gs, ws.XT, rg, pg, 1, 2); }
          else { tile_gemm<1>(de, ws.Gs, ws.XT, rg, pg, 0, 2); tile_gemm<1>(dd, ws.Gs, ws.XT, rg, pg, 1, 2); }
#pragma unroll
          for (int rp = 0; rp < 4; ++rp)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int p = 4 * half + q;
              v[2 * rp][p] = lo_f(yp[rp]) - (lo_f(de[rp][p]) + lo_f(dd[rp][p]));
              v[2 * rp + 1][p] = hi_f(yp[rp]) - (hi_f(de[rp][p]) + hi_f(dd[rp][p]));
            }
        }
        __syncthreads();                           // GEMM2 readers of XT are done
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float* row = ws.XT + (8 * rg + r) * WTP;
          *reinterpret_cast<float4*>(row + 4 * pg) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
          *reinterpret_cast<float4*>(row + 64 + 4 * pg) = make_float4(v[r][4], v[r][5], v[r][6], v[r][7]);
        }
        __syncthreads();
        // ---- phase 3: loglik = c0 - q/2, q = fmaf(v_a, v_a, q) for a ascending (rows a >= ny: v = 0, no-op) ----
        float q = 0.f;
#pragma unroll 8
        for (int a = 0; a < WNX; ++a) {
          const float va = ws.XT[a * WTP + threadIdx.x];
          q = fmaf(va, va, q);
        }
        wv += (double)fmaf(-0.5f, q, Mw.c0);
      }
      if (valid) {
        __stcg(P.w + i, wv);
        acc1.add(wv, dummy, false, sh.mt);
      }
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
    if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0 && P.resampled) P.resampled[k_prop - 1] = res ? 1 : 0;
    sc.t_index += 1;       // filtering.jl:152
  }
  if (k_weigh > 0) {
    const Stats st = reduce_stats<1>(P, sh, acc1, false, cx.red_seq, sc.xseq);
    publish_step<1>(P, sc, k_weigh, st);
  }
}

__global__ void __launch_bounds__(BLOCK, LLPF_MIN_BLOCKS)
k_engine_wide(const __grid_constant__ EngineP P, const __grid_constant__ WideP Mw) {
  __shared__ Shared sh;
  extern __shared__ __align__(16) unsigned char llpf_wide_smem[];
  WideShared& ws = *reinterpret_cast<WideShared*>(llpf_wide_smem);
  math_tab_load(sh.mt);
  for (int k = threadIdx.x; k < WNX * WNX; k += BLOCK) {
    const int c = k / WNX, r = k % WNX;
    ws.As[k] = Mw.At[k];                 // At, Lt are column-major already
    ws.Ls[k] = Mw.Lt[k];
    ws.Gs[k] = Mw.G[r * WNX + c];        // G is row-major: transpose
  }
  if (threadIdx.x < WNX) {
    ws.ldiag[threadIdx.x] = Mw.Lt[threadIdx.x * WNX + threadIdx.x];
    ws.bu[threadIdx.x] = 0.f;
    ws.yt[threadIdx.x] = 0.f;
  }
  __syncthreads();
  Scalars sc = *P.sc;
  Ctx cx;
  cx.bar_target = 0;
  cx.red_seq = 0;
  {
    long long b = (long long)blockIdx.x * P.chunk;
    long long e = b + P.chunk;
    if (b > P.n) b = P.n;
    if (e > P.n) e = P.n;
    cx.beg = (int)b; cx.end = (int)e;
  }
  cx.lwN = -log((double)P.N);
  cx.lw1N = log(1.0 / (double)P.N);
  for (int r = 0; r < P.nops; ++r) {
    const int kind = P.ops[r].kind, a0 = P.ops[r].a0, b0 = P.ops[r].b0, count = P.ops[r].count;
    const int da = P.ops[r].da, db = P.ops[r].db, flags = P.ops[r].flags;
    for (int c = 0; c < count; ++c) {
      if (kind == OP_PF) pf_pass_wide(P, Mw, sh, ws, sc, cx, a0 + c * da, b0 + c * db, flags);
    }
  }
  grid_barrier(P.bar, (unsigned)P.nblocks, cx.bar_target);
  if (blockIdx.x == 0 && threadIdx.x == 0) *P.sc = sc;
}

}  // namespace llpf
