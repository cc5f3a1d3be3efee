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
#include "llpf_engine.cuh"

namespace llpf {

constexpr int WNX = 64;   // padded state dimension of the wide engine (nx <= 64, ny <= 64; padding is zeros)

struct WideP {
  const float* At;        // [64][64]  column-major A:  At[c*64 + r] = A[r,c]
  const float* Lt;        // [64][64]  column-major lower Cholesky factor of R1 (zeros above the diagonal)
  const float* G;         // [64][64]  row-major whitened measurement matrix (rows >= ny are zero)
  const float* B;         // [64][MAX_NU] row-major
  const double* W;        // [ny][ny]  row-major lower: inv(chol(R2))   (yt = W y is formed in f64, then rounded)
  float c0;               // (float) mvnormal_c0
  int nx, ny, nu;
  int diagL;              // L is diagonal: x' += diag(L) z (bit-identical to the general loop: the other terms are +0)
};

// dynamic shared memory of the wide kernel (v2: block-tiled).  The three matrices are stored column by column with every
// element DUPLICATED into a packed f32x2 pair (a, a): the tile GEMMs pair two particles per FFMA2, so the matrix operand is
// the same scalar in both halves and comes straight out of a 16-byte shared load (no register shuffling).
constexpr int WTP = BLOCK;     // particles per tile (one per thread in the load / store / reduce phases)
struct WideShared {
  alignas(16) u64 Ad[WNX * WNX];      // Ad[c*64 + r] = (A[r,c], A[r,c])
  alignas(16) u64 Gd[WNX * WNX];      // Gd[c*64 + a] = (G[a,c], G[a,c])      (rows a >= ny are zero)
  alignas(16) u64 Ld[WNX * WNX];      // Ld[c*64 + r] = (L[r,c], L[r,c])      (lower Cholesky factor of R1; zero above the diagonal)
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
// v2, block-tiled.  v1 gave every thread one particle and its 64 accumulators and streamed all three matrices from shared
// memory for every particle: 16 LDS.128 per 32 FFMA2 — shared-memory bound at 27-29 % of the FP32 peak (profiles/
// r1_v9_ncu_wide.md).  v2 processes a tile of 256 particles per block iteration as two 64 x 64 x 256 GEMMs:
//   thread (warp w, lane l) owns rows 8w..8w+7 of particles {4l..4l+3} and {128+4l..128+4l+3}: 64 accumulators as 32 packed
//   f32x2 pairs (two particles per pair); per column c it loads 8 duplicated matrix elements (4 broadcast LDS.128) and
//   8 particle values (2 conflict-free LDS.128) for 32 FFMA2: 6 LDS.128 instead of 16 per 32 FFMA2.
// Every accumulator still sees exactly the fmaf chain documented at the top of this file (c ascending from 0, then + Bu,
// then the noise chain; even / odd column chains for G x'), so the results are bit-identical to v1 and to the oracle.
// Tile phases (one __syncthreads between them): gather x -> XT | GEMM1 A x (+Bu, +L z) , store x' | x' -> XT | GEMM2 G x',
// v = yt - d | v -> XT | per particle: q = sum v^2 (fmaf chain over a), w += c0 - q/2, online reduction.
__device__ __forceinline__ void lds_4x64(const u64* p, u64 (&v)[8]) {   // 8 consecutive u64 (64 B, 16-byte aligned)
  const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "r"(addr));
  asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[2]), "=l"(v[3]) : "r"(addr + 16));
  asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[4]), "=l"(v[5]) : "r"(addr + 32));
  asm("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[6]), "=l"(v[7]) : "r"(addr + 48));
}
__device__ __forceinline__ void lds_f4_as_pairs(const float* p, u64& a, u64& b) {   // 4 consecutive floats as 2 f32x2 pairs
  lds_2x64(p, a, b);
}

// acc[r][q] += M[8w + r, c] * t[c][particle pair q]   for c = c0, c0 + cs, ... < 64 ; q: pairs 0,1 = particles 4l..4l+3,
// pairs 2,3 = particles 128+4l..128+4l+3 (HALF selects pairs 0..1, 2..3 or all four)
template <int HALF>
__device__ __forceinline__ void tile_gemm(u64 (&acc)[8][4], const u64* Md, const float* XT, int w, int l, int c0, int cs) {
#pragma unroll 2
  for (int c = c0; c < WNX; c += cs) {
    u64 m[8];
    lds_4x64(Md + c * WNX + 8 * w, m);
    u64 x[4];
    if (HALF != 1) lds_f4_as_pairs(XT + c * WTP + 4 * l, x[0], x[1]);
    if (HALF != 0) lds_f4_as_pairs(XT + c * WTP + 128 + 4 * l, x[2], x[3]);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (HALF != 1) { acc[r][0] = ffma2(m[r], x[0], acc[r][0]); acc[r][1] = ffma2(m[r], x[1], acc[r][1]); }
      if (HALF != 0) { acc[r][2] = ffma2(m[r], x[2], acc[r][2]); acc[r][3] = ffma2(m[r], x[3], acc[r][3]); }
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
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
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
    if (prop) {
      // ---- phase 1: x' = A x (+ B u) (+ L z), rows 8w..8w+7 of 8 particles ----
      u64 acc[8][4];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0ull;
      tile_gemm<2>(acc, ws.Ad, ws.XT, w, l, 0, 1);
#pragma unroll
      for (int r = 0; r < 8; ++r) {   // + B u
        const float b = ws.bu[8 * w + r];
        const u64 bb = pack2(b, b);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = fadd2(acc[r][q], bb);
      }
      // noise: 2 Philox blocks (rows 8w..8w+3, 8w+4..8w+7) per particle, f64 Box-Muller rounded to f32 — the same
      // counters as v1: block b of particle gi holds the normals of rows 4b..4b+3
      if (Mw.diagL) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float z[2][8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int pg = P.first + tb + ((q < 2) ? 0 : 128) + 4 * l + 2 * (q & 1) + h;
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
              const uint4 rr = rng_block(P.key, ST_DYN, step_idx, (unsigned long long)(unsigned)pg, (uint32_t)(2 * w + bb));
              const uint32_t ra[2] = {rr.x, rr.z}, rb[2] = {rr.y, rr.w};
              double a0[2], a1[2];
              normal_pairs<2>(ra, rb, a0, a1, sh.mt);
              z[h][4 * bb] = (float)a0[0]; z[h][4 * bb + 1] = (float)a1[0];
              z[h][4 * bb + 2] = (float)a0[1]; z[h][4 * bb + 3] = (float)a1[1];
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float ld = ws.ldiag[8 * w + r];
            acc[r][q] = ffma2(pack2(ld, ld), pack2(z[0][r], z[1][r]), acc[r][q]);
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
        tile_gemm<2>(acc, ws.Ld, ws.XT, w, l, 0, 1);
      }
      // store x' (rows 8w..8w+7 = one 32-byte sector per particle) and hand the tile to phase 2 through XT
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int pl = ((q < 2) ? 0 : 128) + 4 * l + 2 * (q & 1) + h;
          const int ip = tb + pl;
          if (ip < cx.end) {
            float o[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) o[r] = h ? hi_f(acc[r][q]) : lo_f(acc[r][q]);
            stg256(dst + (size_t)ip * WNX + 8 * w, o);
          }
        }
      }
      if (weigh) {
        __syncthreads();                           // readers of XT (x or z) are done
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float* row = ws.XT + (8 * w + r) * WTP;
          const unsigned a0 = (unsigned)__cvta_generic_to_shared(row + 4 * l);
          const unsigned a1 = (unsigned)__cvta_generic_to_shared(row + 128 + 4 * l);
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a0), "l"(acc[r][0]), "l"(acc[r][1]) : "memory");
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a1), "l"(acc[r][2]), "l"(acc[r][3]) : "memory");
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
        float v[8][8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          u64 de[8][4], dd[8][4];
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) { de[r][q] = 0ull; dd[r][q] = 0ull; }
          if (half == 0) { tile_gemm<0>(de, ws.Gd, ws.XT, w, l, 0, 2); tile_gemm<0>(dd, ws.Gd, ws.XT, w, l, 1, 2); }
          else { tile_gemm<1>(de, ws.Gd, ws.XT, w, l, 0, 2); tile_gemm<1>(dd, ws.Gd, ws.XT, w, l, 1, 2); }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float y = ws.yt[8 * w + r];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const u64 e = de[r][2 * half + q], o = dd[r][2 * half + q];
              v[r][4 * half + 2 * q] = y - (lo_f(e) + lo_f(o));
              v[r][4 * half + 2 * q + 1] = y - (hi_f(e) + hi_f(o));
            }
          }
        }
        __syncthreads();                           // GEMM2 readers of XT are done
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float* row = ws.XT + (8 * w + r) * WTP;
          *reinterpret_cast<float4*>(row + 4 * l) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
          *reinterpret_cast<float4*>(row + 128 + 4 * l) = make_float4(v[r][4], v[r][5], v[r][6], v[r][7]);
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

__global__ void __launch_bounds__(BLOCK, 1)
k_engine_wide(const __grid_constant__ EngineP P, const __grid_constant__ WideP Mw) {
  __shared__ Shared sh;
  extern __shared__ __align__(16) unsigned char llpf_wide_smem[];
  WideShared& ws = *reinterpret_cast<WideShared*>(llpf_wide_smem);
  math_tab_load(sh.mt);
  for (int k = threadIdx.x; k < WNX * WNX; k += BLOCK) {
    const int c = k / WNX, r = k % WNX;
    const float a = Mw.At[k], lv = Mw.Lt[k], g = Mw.G[r * WNX + c];   // At, Lt column-major; G row-major
    ws.Ad[k] = pack2(a, a);
    ws.Ld[k] = pack2(lv, lv);
    ws.Gd[k] = pack2(g, g);
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

// reset!(pf)  filtering.jl:4-14 for wide models: x0 = mu0 + L0 z  (f32: fmaf chain from 0 over c <= r, then + mu0)
__global__ void k_init_wide(float* x, long long n, long long first, RngKey key, const float* mu0, const float* L0 /*row-major*/,
                            int nx) {
  __shared__ MathTab mt;
  math_tab_load(mt);
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float z[WNX];
#pragma unroll 1
    for (int b = 0; b < WNX / 4; ++b) {
      const uint4 r = rng_block(key, ST_INIT, 0u, (unsigned long long)(first + i), (uint32_t)b);
      const uint32_t ra[2] = {r.x, r.z}, rb[2] = {r.y, r.w};
      double a0[2], a1[2];
      normal_pairs<2>(ra, rb, a0, a1, mt);
      z[4 * b] = (float)a0[0]; z[4 * b + 1] = (float)a1[0]; z[4 * b + 2] = (float)a0[1]; z[4 * b + 3] = (float)a1[1];
    }
    float* xo = x + (size_t)i * WNX;
    for (int r = 0; r < WNX; ++r) {
      float acc = 0.f;
      if (r < nx) {
        for (int c = 0; c <= r; ++c) acc = fmaf(L0[r * WNX + c], z[c], acc);
        acc = mu0[r] + acc;
      }
      xo[r] = acc;
    }
  }
}

__global__ void k_export_x_wide(const float* x, long long n, int nx, double* out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    for (int d = 0; d < nx; ++d) out[(size_t)i * nx + d] = (double)x[(size_t)i * WNX + d];
}
__global__ void k_import_x_wide(float* x, long long n, int nx, const double* in) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    for (int d = 0; d < WNX; ++d) x[(size_t)i * WNX + d] = d < nx ? (float)in[(size_t)i * nx + d] : 0.f;
}
// (sum we, sum we^2, sum we*x[d]) block partials for weighted_mean / effective_particles  (filtering.jl:541-568)
__global__ void k_wstats_wide(const double* we, const float* x, long long n, int nx, double* part /*[grid][2+WNX]*/) {
  __shared__ double sm[8 * (2 + WNX)];
  double v[2 + WNX];
  for (int k = 0; k < 2 + WNX; ++k) v[k] = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double e = we[i];
    v[0] += e;
    v[1] = fma(e, e, v[1]);
    for (int d = 0; d < nx; ++d) v[2 + d] = fma(e, (double)x[(size_t)i * WNX + d], v[2 + d]);
  }
  for (int k = 0; k < 2 + WNX; ++k)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 2 + WNX; ++k) sm[(threadIdx.x >> 5) * (2 + WNX) + k] = v[k];
  __syncthreads();
  if (threadIdx.x < 2 + WNX) {
    double r = 0.0;
    for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) r += sm[wq * (2 + WNX) + threadIdx.x];
    part[(size_t)blockIdx.x * (2 + WNX) + threadIdx.x] = r;
  }
}

}  // namespace llpf
