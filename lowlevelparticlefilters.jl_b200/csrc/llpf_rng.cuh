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
#include <cuda_runtime.h>
#include <stdint.h>

namespace llpf {

enum : uint32_t { ST_INIT = 0, ST_DYN = 1, ST_RESAMPLE = 2, ST_STRAT = 3, ST_RESID = 4 };

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

__device__ __forceinline__ void normal_pair(uint32_t ra, uint32_t rb, double& z0, double& z1) {
  const double u1 = uniform32_open(ra);
  // 2*u2 = (rb + 0.5) * 2^-31, exact
  const double a2 = fma((double)rb, 4.6566128730773926e-10, 2.3283064365386963e-10);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(a2, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

// N standard normals for (stream, step, particle i), blocks of 4 per Philox call
template <int N>
__device__ __forceinline__ void normals(const RngKey& key, uint32_t stream, uint32_t step,
                                        unsigned long long i, double (&z)[N]) {
#pragma unroll
  for (int b = 0; 4 * b < N; ++b) {
    const uint4 r = rng_block(key, stream, step, i, (uint32_t)b);
    double a0, a1, a2, a3;
    normal_pair(r.x, r.y, a0, a1);
    z[4 * b] = a0;
    if (4 * b + 1 < N) z[4 * b + 1] = a1;
    if (4 * b + 2 < N) {
      normal_pair(r.z, r.w, a2, a3);
      z[4 * b + 2] = a2;
      if (4 * b + 3 < N) z[4 * b + 3] = a3;
    }
  }
}

}  // namespace llpf
