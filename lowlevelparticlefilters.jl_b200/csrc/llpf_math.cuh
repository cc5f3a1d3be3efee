// llpf_math.cuh — domain-specialised f64 math for the particle loop.
//
// The particle sweep needs exactly four transcendental shapes, each on a known, benign domain:
//   ln(u)        u = (r + 0.5) 2^-32, r a 32-bit integer            (Box-Muller radius)
//   sqrt(x)      x = -2 ln(u) in [2.3e-10, 46]
//   sin/cos(pi a) a = (r + 0.5) 2^-31                                 (Box-Muller angle)
//   exp(x)       x <= 0                                               (weights relative to a maximum)
// CUDA's general-purpose log/sqrt/sincospi/exp spend most of their instructions on argument
// classification, special cases and 64-bit immediates (materialised through UMOV pairs).  Here the
// reductions start from the integer, there are no branches, tables live in shared memory and the
// polynomial coefficients in constant memory (DFMA takes c[bank][offset] operands directly).
// Accuracy: <= ~1.5 ulp (validated on the device against the CPU oracle's libm, tests/test_gpu_parity.py).
#pragma once
#include "llpf_rtc_compat.h"

namespace llpf {

#include "llpf_math_tables.inc"

struct MathTab {
  double2 lg[129];   // (1/c_k rounded, -ln of that), c_k = 1 + k/128 (k < 64), (1 + k/128)/2 (k >= 64)
  double ex[64];     // 2^(j/64)
  // polynomial coefficients, read with volatile 16-byte shared loads inside the particle loop: one
  // LDS.128 per two constants, and — unlike __constant__/immediate operands — nothing the compiler can
  // hoist out of the loop into (spilled) registers
  double2 sc[9];     // (sinpi_i, cospi_i)
  double2 l1p[3];    // log1p: (c0,c1) (c2,c3) (c4,c5)
  double2 ln2;       // (hi, lo)
  double2 ep[3];     // exp: (p0,p1) (p2,p3) (p4, 64/ln2)
  double2 eh;        // ln2/64 (hi, lo)
};

// cooperative load of the tables into shared memory (call once per block, then __syncthreads)
__device__ __forceinline__ void math_tab_load(MathTab& T) {
  for (int k = threadIdx.x; k < 129; k += blockDim.x) T.lg[k] = make_double2(c_log_tab[2 * k], c_log_tab[2 * k + 1]);
  for (int k = threadIdx.x; k < 64; k += blockDim.x) T.ex[k] = c_exp_tab[k];
  if (threadIdx.x < 9) T.sc[threadIdx.x] = make_double2(c_sinpi[threadIdx.x], c_cospi[threadIdx.x]);
  if (threadIdx.x < 3) T.l1p[threadIdx.x] = make_double2(c_l1p[2 * threadIdx.x], c_l1p[2 * threadIdx.x + 1]);
  if (threadIdx.x == 0) {
    T.ln2 = make_double2(c_ln2_hi, c_ln2_lo);
    T.ep[0] = make_double2(c_exp_p[0], c_exp_p[1]);
    T.ep[1] = make_double2(c_exp_p[2], c_exp_p[3]);
    T.ep[2] = make_double2(c_exp_p[4], c_exp_inv);
    T.eh = make_double2(c_exp_hi, c_exp_lo);
  }
}

// volatile 16-byte shared load (not hoistable, not mergeable)
__device__ __forceinline__ double2 lds2v(const double2* p) {
  double2 v;
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}

// ln((r + 0.5) * 2^-32) for V independent 32-bit words (constants fetched once for all V chains)
template <int V>
__device__ __forceinline__ void log_u32_v(const uint32_t (&r)[V], double (&L)[V], const MathTab& T) {
  double rr[V], ef[V], lc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double d = fma((double)r[v], 2.0, 1.0);          // 2r+1, exact (33 bits)
    const int hi = __double2hiint(d), lo = __double2loint(d);
    int e = (hi >> 20) - (1023 + 33);                      // u = m * 2^e, m in [1,2)
    const int mh = hi & 0x000fffff;
    const int k = (mh + 0x1000) >> 13;                     // round((m-1)*128) in [0,128]
    const int wrap = (k + 64) >> 7;                        // m >= 1.5: use m/2, e+1 (u in [0.75,1) gets e = 0:
    e += wrap;                                             //  no e*ln2 term, so ln(u -> 1) keeps full relative accuracy)
    const double m = __hiloint2double(mh | (0x3ff00000 - (wrap << 20)), lo);
    const double2 t = T.lg[k];
    rr[v] = fma(m, t.x, -1.0);                             // |rr| <= 2^-8
    lc[v] = t.y;
    ef[v] = (double)e;
  }
  const double2 c45 = lds2v(&T.l1p[2]), c23 = lds2v(&T.l1p[1]), c01 = lds2v(&T.l1p[0]), ln2 = lds2v(&T.ln2);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    double p = fma(rr[v], c45.y, c45.x);
    p = fma(rr[v], p, c23.y);
    p = fma(rr[v], p, c23.x);
    p = fma(rr[v], p, c01.y);
    p = fma(rr[v], p, c01.x);
    const double l1p = fma(rr[v] * rr[v], p, rr[v]);       // log1p(rr)
    const double hi_part = fma(ef[v], ln2.x, lc[v]);       // e*ln2_hi is exact
    L[v] = hi_part + fma(ef[v], ln2.y, l1p);
  }
}

// sqrt(x), x normal positive (no zero / denormal / inf handling): rsqrt seed + one coupled Newton step + correction
__device__ __forceinline__ double sqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y;
  double h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double dd = fma(-g, g, x);
  return fma(dd, h, g);
}

// (sin, cos)(pi * (r + 0.5) * 2^-31), r in [0, 2^32): quadrant from the top bits, Taylor on |t| <= 1/4
template <int V>
__device__ __forceinline__ void sincospi_u32_v(const uint32_t (&r)[V], double (&s)[V], double (&c)[V], const MathTab& T) {
  double t[V], t2[V], sp[V], cp[V];
  uint32_t k[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    k[v] = ((r[v] >> 29) + 1u) >> 1;                       // round(2a) in [0,4]
    const int ti = (int)(r[v] - (k[v] << 30));             // wraps mod 2^32: in [-2^29, 2^29)
    t[v] = fma((double)ti, 4.6566128730773926e-10, 2.3283064365386963e-10);  // (ti + 0.5) 2^-31, exact
    t2[v] = t[v] * t[v];
  }
#ifndef LLPF_SC_BATCH
#define LLPF_SC_BATCH 3
#endif
  // coefficients are fetched LLPF_SC_BATCH at a time so that their LDS.128 latencies overlap
  {
    const double2 c8 = lds2v(&T.sc[8]), c7 = lds2v(&T.sc[7]);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      sp[v] = fma(t2[v], c8.x, c7.x);
      cp[v] = fma(t2[v], c8.y, c7.y);
    }
  }
#pragma unroll
  for (int i0 = 6; i0 >= 0; i0 -= LLPF_SC_BATCH) {
    double2 co[LLPF_SC_BATCH];
#pragma unroll
    for (int b = 0; b < LLPF_SC_BATCH; ++b)
      if (i0 - b >= 0) co[b] = lds2v(&T.sc[i0 - b]);
#pragma unroll
    for (int b = 0; b < LLPF_SC_BATCH; ++b) {
      if (i0 - b >= 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          sp[v] = fma(t2[v], sp[v], co[b].x);
          cp[v] = fma(t2[v], cp[v], co[b].y);
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double st = t[v] * sp[v];
    const bool swap = (k[v] & 1u) != 0;
    const double ss = swap ? cp[v] : st;
    const double cc = swap ? st : cp[v];
    const int sneg = (int)((k[v] & 2u) << 30);             // sign bit if k in {2,3}
    const int cneg = (int)(((k[v] + 1u) & 2u) << 30);      // sign bit if k in {1,2}
    s[v] = __hiloint2double(__double2hiint(ss) ^ sneg, __double2loint(ss));
    c[v] = __hiloint2double(__double2hiint(cc) ^ cneg, __double2loint(cc));
  }
}

// exp(x) for x <= 0 (x = -inf allowed -> 0; results below ~1e-307 flush to 0)
__device__ __forceinline__ double exp_nonpos(double x, const MathTab& T) {
  const double magic = 6755399441055744.0;               // 1.5 * 2^52: rint via add
  const double2 p01 = lds2v(&T.ep[0]), p23 = lds2v(&T.ep[1]), p4i = lds2v(&T.ep[2]), hl = lds2v(&T.eh);
  const double xs = fmax(x, -720.0);
  const double nm = fma(xs, p4i.y, magic);
  const int n = __double2loint(nm);                      // rint(x * 64/ln2)
  const double nf = nm - magic;
  double r = fma(nf, -hl.x, xs);
  r = fma(nf, -hl.y, r);                                 // |r| <= ln2/128
  double p = fma(r, p4i.x, p23.y);
  p = fma(r, p, p23.x);
  p = fma(r, p, p01.y);
  p = fma(r, p, p01.x);
  const double q = fma(r * r, p, r);                     // exp(r) - 1
  const double tj = T.ex[n & 63];
  const double v = fma(tj, q, tj);                       // 2^(j/64) * exp(r) in [1, 2)
  const int sh = n >> 6;                                 // power of two, in [-1039, 0]
  const double res = __hiloint2double(__double2hiint(v) + (sh << 20), __double2loint(v));
  return (x < -708.0) ? 0.0 : res;
}

}  // namespace llpf
