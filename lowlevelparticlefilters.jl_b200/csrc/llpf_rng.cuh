// llpf_rng.cuh — device side of the RNG contract (DESIGN.md "RNG contract").
//
// The reference draws randn() from a sequential Xoshiro (src/PFtypes.jl:30,135,153) and rand() from
// the task-local global RNG (src/resample.jl:23,49); neither can be replayed by 10^6 threads.  The
// contract replaces them by a stateless counter-based generator so that a particle's variates depend
// only on (seed, epoch, stream, step, global particle index) — identical on 1 or 8 GPUs and in the
// CPU oracle, which implements the same contract independently (oracle/llpf_oracle.c).
//
//   key     = (seed[31:0], seed[63:32])
//   counter = (i[31:0], blk + (i[63:32] << 16), step, stream | epoch << 8)
//   Philox4x32-10 (Salmon et al., SC'11; checked against the Random123 known-answer vectors)
//   uniform32: u = (r + 0.5) * 2^-32  in (0,1), exact in f64
//   uniform53: u = (((r0 << 32) | r1) >> 11) * 2^-53 in [0,1)   (granularity of Julia's rand())
//   normal pair: Box-Muller in f64: rad = sqrt(-2 log u1); (z0,z1) = rad * (cospi(2 u2), sinpi(2 u2))
#pragma once
#include "llpf_rtc_compat.h"

#include "llpf_math.cuh"

namespace llpf {

enum : uint32_t { ST_INIT = 0, ST_DYN = 1, ST_RESAMPLE = 2, ST_STRAT = 3, ST_RESID = 4, ST_SMOOTH = 6 };   // 5: data simulation (oracle)

struct RngKey {
  uint32_t k0, k1;    // seed
  uint32_t epoch8;    // epoch << 8
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 rng_block(const RngKey& key, uint32_t stream, uint32_t step,
                                           unsigned long long i, uint32_t blk) {
  return philox4x32_10((uint32_t)i, blk + ((uint32_t)(i >> 32) << 16), step, stream | key.epoch8,
                       key.k0, key.k1);
}

__device__ __forceinline__ double uniform53(uint32_t hi, uint32_t lo) {
  const unsigned long long v = (((unsigned long long)hi << 32) | lo) >> 11;
  return (double)v * 1.1102230246251565e-16;
}

// (r + 0.5) * 2^-32, exact
__device__ __forceinline__ double uniform32_open(uint32_t r) {
  return fma((double)r, 2.3283064365386963e-10, 1.1641532182693481e-10);
}

// Box-Muller on V pairs of 32-bit words at once (constants of llpf_math.cuh fetched once per call):
// z0 = rad cos(2 pi u2), z1 = rad sin(2 pi u2), rad = sqrt(-2 ln u1), u1 = (ra+0.5) 2^-32, 2 u2 = (rb+0.5) 2^-31
template <int V>
__device__ __forceinline__ void normal_pairs(const uint32_t (&ra)[V], const uint32_t (&rb)[V], double (&z0)[V],
                                             double (&z1)[V], const MathTab& T) {
  double L[V], s[V], c[V];
  log_u32_v<V>(ra, L, T);
  sincospi_u32_v<V>(rb, s, c, T);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double rad = sqrt_pos(-2.0 * L[v]);
    z0[v] = rad * c[v];
    z1[v] = rad * s[v];
  }
}

// N standard normals for (stream, step, particle i), blocks of 4 per Philox call
template <int N>
__device__ __forceinline__ void normals(const RngKey& key, uint32_t stream, uint32_t step,
                                        unsigned long long i, double (&z)[N], const MathTab& T) {
#pragma unroll
  for (int b = 0; 4 * b < N; ++b) {
    const uint4 r = rng_block(key, stream, step, i, (uint32_t)b);
    if (4 * b + 2 < N) {
      const uint32_t ra[2] = {r.x, r.z}, rb[2] = {r.y, r.w};
      double a0[2], a1[2];
      normal_pairs<2>(ra, rb, a0, a1, T);
      z[4 * b] = a0[0];
      z[4 * b + 1] = a1[0];
      z[4 * b + 2] = a0[1];
      if (4 * b + 3 < N) z[4 * b + 3] = a1[1];
    } else {
      const uint32_t ra[1] = {r.x}, rb[1] = {r.y};
      double a0[1], a1[1];
      normal_pairs<1>(ra, rb, a0, a1, T);
      z[4 * b] = a0[0];
      if (4 * b + 1 < N) z[4 * b + 1] = a1[0];
    }
  }
}

}  // namespace llpf
