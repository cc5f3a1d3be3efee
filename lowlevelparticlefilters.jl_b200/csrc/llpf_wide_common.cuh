// llpf_wide_common.cuh — what the host API (llpf_api.cu) and the wide engine (llpf_wide.cu) share: the model descriptor of
// the Float32-particle engine and the small utility kernels (reset!, AoS import / export, weighted statistics).
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

// the engine itself lives in its own translation unit (llpf_wide.cu, compiled with its own block size)
const void* wide_engine_kernel();
size_t wide_engine_smem_bytes();
int wide_engine_block_threads();
int wide_engine_min_blocks();

#ifndef LLPF_WIDE_ENGINE_TU   // the utility kernels belong to the host API translation unit only
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

#endif  // LLPF_WIDE_ENGINE_TU

}  // namespace llpf
