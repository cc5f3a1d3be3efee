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

// dynamic shared memory of the wide kernel
struct WideShared {
  alignas(16) float As[WNX * WNX];
  alignas(16) float Gs[WNX * WNX];
  alignas(16) float Ls[WNX * WNX];
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

// acc[0..63] += column c of a column-major 64x64 matrix in shared memory, times v:  32 FFMA2 + 16 LDS.128
__device__ __forceinline__ void axpy_col64(u64 (&acc)[WNX / 2], const float* col, float v) {
  const u64 vv = pack2(v, v);
#pragma unroll
  for (int r4 = 0; r4 < WNX / 4; ++r4) {
    u64 a01, a23;
    lds_2x64(col + 4 * r4, a01, a23);
    acc[2 * r4] = ffma2(a01, vv, acc[2 * r4]);
    acc[2 * r4 + 1] = ffma2(a23, vv, acc[2 * r4 + 1]);
  }
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
  const int ny4 = (Mw.ny + 3) & ~3;
  Online<1> acc1;
  acc1.init();
  const double dummy[1] = {0.0};
  for (int i = cx.beg + threadIdx.x; i < cx.end; i += BLOCK) {
    const int gi = P.first + i;
    // ancestor (resample.jl:26-34: slots past the last threshold keep state.j)
    int a = gi;
    if (res) {
      a = __ldcg(P.j + i);
      if (gi >= f_total) {
        if (jid) a = gi;
        __stcg(P.j + i, a);
      }
    }
    double wraw = 0.0;
    if (!res && !wst.uniform) wraw = __ldcg(P.w + i);
    u64 acc[WNX / 2];
    if (k_prop > 0) {
      const float* xin;
      if (P.world > 1 && a < 0) {   // an ancestor another rank shipped here as a packed entry (expand_packs)
        const char* e = P.pack_in + (size_t)(-1 - a) * (size_t)P.pack_stride;
        xin = reinterpret_cast<const float*>(e);
        __stcg(P.j + i, __ldcg(reinterpret_cast<const int*>(e + P.pack_state_bytes)));   // state.j keeps global ids
      } else {
        xin = wide_row(P, cur, a);
      }
#pragma unroll
      for (int k = 0; k < WNX / 2; ++k) acc[k] = 0ull;
      // x' = A x : column form, x streamed 8 values (one 32-byte sector) at a time
#pragma unroll 1
      for (int c8 = 0; c8 < WNX / 8; ++c8) {
        float xv[8];
        ldg256(xin + 8 * c8, xv);
#pragma unroll
        for (int k = 0; k < 8; ++k) axpy_col64(acc, ws.As + (8 * c8 + k) * WNX, xv[k]);
      }
      // + B u
#pragma unroll
      for (int r4 = 0; r4 < WNX / 4; ++r4) {
        u64 b01, b23;
        lds_2x64(ws.bu + 4 * r4, b01, b23);
        acc[2 * r4] = fadd2(acc[2 * r4], b01);
        acc[2 * r4 + 1] = fadd2(acc[2 * r4 + 1], b23);
      }
      // + L z : 16 Philox blocks of 4 normals (f64 Box-Muller, rounded to f32)
      if (Mw.diagL) {
#pragma unroll 1
        for (int b = 0; b < WNX / 4; ++b) {
          const uint4 r = rng_block(P.key, ST_DYN, step_idx, (unsigned long long)(unsigned)gi, (uint32_t)b);
          const uint32_t ra[2] = {r.x, r.z}, rb[2] = {r.y, r.w};
          double a0[2], a1[2];
          normal_pairs<2>(ra, rb, a0, a1, sh.mt);
          u64 l01, l23;
          lds_2x64(ws.ldiag + 4 * b, l01, l23);
          const u64 z01 = pack2((float)a0[0], (float)a1[0]), z23 = pack2((float)a0[1], (float)a1[1]);
          // acc[2b], acc[2b+1] with a runtime index would force the array into local memory: select statically
#pragma unroll
          for (int q = 0; q < WNX / 4; ++q) {
            if (q == b) {
              acc[2 * q] = ffma2(l01, z01, acc[2 * q]);
              acc[2 * q + 1] = ffma2(l23, z23, acc[2 * q + 1]);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int b = 0; b < WNX / 4; ++b) {
          const uint4 r = rng_block(P.key, ST_DYN, step_idx, (unsigned long long)(unsigned)gi, (uint32_t)b);
          const uint32_t ra[2] = {r.x, r.z}, rb[2] = {r.y, r.w};
          double a0[2], a1[2];
          normal_pairs<2>(ra, rb, a0, a1, sh.mt);
          const float z[4] = {(float)a0[0], (float)a1[0], (float)a0[1], (float)a1[1]};
#pragma unroll
          for (int k = 0; k < 4; ++k) axpy_col64(acc, ws.Ls + (4 * b + k) * WNX, z[k]);
        }
      }
      float* xo = dst + (size_t)i * WNX;
#pragma unroll
      for (int c8 = 0; c8 < WNX / 8; ++c8) {
        const float o[8] = {lo_f(acc[4 * c8]),     hi_f(acc[4 * c8]),     lo_f(acc[4 * c8 + 1]), hi_f(acc[4 * c8 + 1]),
                            lo_f(acc[4 * c8 + 2]), hi_f(acc[4 * c8 + 2]), lo_f(acc[4 * c8 + 3]), hi_f(acc[4 * c8 + 3])};
        stg256(xo + 8 * c8, o);
      }
    } else if (k_weigh > 0) {
      // correct! only: the particle as it is
      const float* xin = wide_row(P, cur, gi);
#pragma unroll
      for (int c8 = 0; c8 < WNX / 8; ++c8) {
        float xv[8];
        ldg256(xin + 8 * c8, xv);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[4 * c8 + k] = pack2(xv[2 * k], xv[2 * k + 1]);
      }
    }
    double wv;
    if (res) wv = cx.lw1N;                    // reset_weights!  utils.jl:75
    else wv = wst.uniform ? wst.wu : (wst.pend ? (wraw - wst.pm) - wst.pls : wraw);
    if (k_weigh > 0) {
      if (!skip) {
        // loglik = c0 - |yt - G x'|^2 / 2 ; four rows of G at a time, each row two fmaf chains (even / odd columns)
        float q = 0.f;
#pragma unroll 1
        for (int a0 = 0; a0 < ny4; a0 += 4) {
          u64 d[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
          for (int c4 = 0; c4 < WNX / 4; ++c4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              u64 g01, g23;
              lds_2x64(ws.Gs + (a0 + k) * WNX + 4 * c4, g01, g23);
              d[k] = ffma2(g01, acc[2 * c4], d[k]);
              d[k] = ffma2(g23, acc[2 * c4 + 1], d[k]);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float v = ws.yt[a0 + k] - (lo_f(d[k]) + hi_f(d[k]));
            q = fmaf(v, v, q);
          }
        }
        wv += (double)fmaf(-0.5f, q, Mw.c0);
      }
      __stcg(P.w + i, wv);
      acc1.add(wv, dummy, false, sh.mt);
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
    ws.As[k] = Mw.At[k];
    ws.Gs[k] = Mw.G[k];
    ws.Ls[k] = Mw.Lt[k];
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
