/*
 * llpf_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the particle-filter hot path of
 * baggepinnen/LowLevelParticleFilters.jl (commit d5396f6, v3.31.1).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may call it; the
 * product (libllpf_b200.so) never links, loads or falls back to anything in this directory.
 *
 * PARITY STATUS: "parity unpinned" at the trajectory level.  The reference is pure Julia and no
 * julia binary exists in the build container, so this file could not be validated against
 * outputs of the reference itself, and the reference ships no golden trajectories for this path.
 * What it IS pinned against (tests/test_oracle_kat.py): every known-answer / invariant test the
 * reference's own test-suite holds for the path — test/runtests.jl:29-47 (logsumexp!/expnormalize!),
 * :88-106 (systematic resample known answers), :108-143 (resample proportions), :145-154
 * (stratified known answer), :182-188 (rk4), :412-449 (PF/APF log-likelihood vs Kalman filter) —
 * plus the closed-form Kalman filter (src/filtering.jl:52-133) as an RNG-independent check.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 * Arithmetic is written in the reference's evaluation order; compile with -ffp-contract=off so gcc
 * does not fuse what Julia does not fuse.
 *
 * Randomness: the reference draws randn from a sequential Xoshiro and rand() from the task-local
 * global RNG (src/resample.jl:23,49,106), which no parallel implementation can replay.  "Identical
 * RNG streams" is therefore defined by the counter-based contract in DESIGN.md ("RNG contract"):
 * Philox4x32-10 keyed by the seed, counter = (particle, block, step, stream|epoch<<8), 32-bit
 * uniforms -> Box-Muller in f64.  This file implements that contract independently of the CUDA code.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/llpf.h"

/* ------------------------------------------------------------------------------------------
 * RNG contract (ours, not the reference's)
 * ---------------------------------------------------------------------------------------- */
enum { STREAM_INIT = 0, STREAM_DYN = 1, STREAM_RESAMPLE = 2, STREAM_STRAT = 3, STREAM_RESID = 4,
       STREAM_SIM = 5, STREAM_SMOOTH = 6 };

static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k[0];
  const uint32_t n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}

static void rng_block(uint64_t seed, uint64_t epoch, uint32_t stream, uint32_t t, uint64_t i,
                      uint32_t blk, uint32_t out[4]) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {(uint32_t)i, blk + ((uint32_t)(i >> 32) << 16), t,
                     stream | ((uint32_t)epoch << 8)};
  orc_philox4x32_10(ctr, key, out);
}

/* u in (0,1), exactly representable: (r + 0.5) * 2^-32 */
static inline double u32_open(uint32_t r) { return ((double)r + 0.5) * 2.3283064365386963e-10; }
/* u in [0,1), 53 bits, the granularity of Julia's rand(Float64) */
static inline double u53(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * 1.1102230246251565e-16;
}
double orc_uniform53(uint64_t seed, uint64_t epoch, uint32_t stream, uint32_t t, uint64_t i) {
  uint32_t r[4];
  rng_block(seed, epoch, stream, t, i, 0, r);
  return u53(r[0], r[1]);
}

/* sin(pi a), cos(pi a) for a in (0,2): exact quadrant reduction, then libm on [-pi/4, pi/4] */
static void sincospi_02(double a, double* s, double* c) {
  const double q = nearbyint(2.0 * a);       /* 0..4 */
  const double r = a - 0.5 * q;              /* exact, |r| <= 0.25 */
  const double sr = sin(M_PI * r), cr = cos(M_PI * r);
  switch (((int)q) & 3) {
    case 0: *s = sr;  *c = cr;  break;
    case 1: *s = cr;  *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}

/* Box-Muller: two 32-bit words -> two N(0,1) */
static void normal_pair(uint32_t ra, uint32_t rb, double* z0, double* z1) {
  const double u1 = u32_open(ra);
  const double u2 = u32_open(rb);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi_02(2.0 * u2, &s, &c);
  *z0 = rad * c;
  *z1 = rad * s;
}

/* n standard normals for (stream, t, particle i) */
void orc_normals(uint64_t seed, uint64_t epoch, uint32_t stream, uint32_t t, uint64_t i, int n,
                 double* z) {
  for (int b = 0; 4 * b < n; ++b) {
    uint32_t r[4];
    double zz[4];
    rng_block(seed, epoch, stream, t, i, (uint32_t)b, r);
    normal_pair(r[0], r[1], &zz[0], &zz[1]);
    normal_pair(r[2], r[3], &zz[2], &zz[3]);
    for (int k = 0; k < 4 && 4 * b + k < n; ++k) z[4 * b + k] = zz[k];
  }
}

/* ------------------------------------------------------------------------------------------
 * small dense helpers (column-major)
 * ---------------------------------------------------------------------------------------- */
#define CM(M, r, c, ld) ((M)[(size_t)(c) * (ld) + (r)])

/* lower Cholesky, returns 0 on success. L is n*n column-major, upper part zeroed. */
int orc_cholesky_lower(const double* S, int n, double* L) {
  memset(L, 0, sizeof(double) * n * n);
  for (int jc = 0; jc < n; ++jc) {
    double d = CM(S, jc, jc, n);
    for (int k = 0; k < jc; ++k) d -= CM(L, jc, k, n) * CM(L, jc, k, n);
    if (!(d > 0.0)) return 1;
    d = sqrt(d);
    CM(L, jc, jc, n) = d;
    for (int i = jc + 1; i < n; ++i) {
      double v = CM(S, i, jc, n);
      for (int k = 0; k < jc; ++k) v -= CM(L, i, k, n) * CM(L, jc, k, n);
      CM(L, i, jc, n) = v / d;
    }
  }
  return 0;
}

/* out = M*v, StaticArrays order: out[r] = M[r,1]*v[1] + M[r,2]*v[2] + ... left to right */
static void matvec(const double* M, int rows, int cols, const double* v, double* out) {
  for (int r = 0; r < rows; ++r) {
    double acc = CM(M, r, 0, rows) * v[0];
    for (int c = 1; c < cols; ++c) acc = acc + CM(M, r, c, rows) * v[c];
    out[r] = acc;
  }
}

/* Julia Base mapreduce pairwise sum (base/reduce.jl mapreduce_impl, blksize 1024): sequential up to
   1024 elements, else split at ifirst + (ilast-ifirst)>>1.  (Inside a block Julia's @simd may
   re-associate; that is compiler-dependent and not reproducible — sequential is the restatement.) */
static double pairwise_sum(const double* a, int64_t n, int square) {
  if (n <= 0) return 0.0;
  if (n <= 1024) {
    double s = square ? a[0] * a[0] : a[0];
    for (int64_t i = 1; i < n; ++i) s += square ? a[i] * a[i] : a[i];
    return s;
  }
  const int64_t h = ((n - 1) >> 1) + 1;
  return pairwise_sum(a, h, square) + pairwise_sum(a + h, n - h, square);
}
double orc_sum(const double* a, int64_t n) { return pairwise_sum(a, n, 0); }

/* ------------------------------------------------------------------------------------------
 * weight numerics  (src/utils.jl)
 * ---------------------------------------------------------------------------------------- */
/* findmax: value and FIRST index of the maximum (Julia findmax) */
static double findmax(const double* w, int64_t n, int64_t* ind) {
  double m = w[0];
  int64_t k = 0;
  for (int64_t i = 1; i < n; ++i)
    if (w[i] > m) { m = w[i]; k = i; }
  *ind = k;
  return m;
}

/* sum_all_but(w,i)  utils.jl:66-71 */
static double sum_all_but(double* we, int64_t n, int64_t i) {
  we[i] -= 1;
  const double s = pairwise_sum(we, n, 0);
  we[i] += 1;
  return s;
}

/* logsumexp!(w,we,maxw)  utils.jl:18-27 (exp_map! :3-7 uses SLEEFPirates.exp, <=1 ulp vs libm) */
double orc_logsumexp(double* w, double* we, int64_t n, double* maxw) {
  int64_t maxind;
  const double offset = findmax(w, n, &maxind);
  for (int64_t i = 0; i < n; ++i) w[i] -= offset;
  for (int64_t i = 0; i < n; ++i) we[i] = exp(w[i]);
  const double s = sum_all_but(we, n, maxind);
  const double inv = 1 / (s + 1);
  for (int64_t i = 0; i < n; ++i) we[i] *= inv;
  const double l1p = log1p(s);
  for (int64_t i = 0; i < n; ++i) w[i] -= l1p;
  if (maxw) *maxw = offset;
  return l1p + offset;
}

/* expnormalize!(we,w)  utils.jl:48-55 : w unchanged (up to w-offset+offset rounding) */
void orc_expnormalize2(double* we, double* w, int64_t n) {
  int64_t maxind;
  const double offset = findmax(w, n, &maxind);
  for (int64_t i = 0; i < n; ++i) w[i] -= offset;
  for (int64_t i = 0; i < n; ++i) we[i] = exp(w[i]);
  for (int64_t i = 0; i < n; ++i) w[i] += offset;
  const double s = sum_all_but(we, n, maxind);
  const double inv = 1 / (s + 1);
  for (int64_t i = 0; i < n; ++i) we[i] *= inv;
}

/* expnormalize!(w)  utils.jl:57-63 : in place */
void orc_expnormalize1(double* w, int64_t n) {
  int64_t maxind;
  const double offset = findmax(w, n, &maxind);
  for (int64_t i = 0; i < n; ++i) w[i] -= offset;
  for (int64_t i = 0; i < n; ++i) w[i] = exp(w[i]);
  const double s = sum_all_but(w, n, maxind);
  const double inv = 1 / (s + 1);
  for (int64_t i = 0; i < n; ++i) w[i] *= inv;
}

/* effective_particles(we) = 1/sum(abs2, we)   resample.jl:1-2 */
double orc_effective_particles(const double* we, int64_t n) { return 1 / pairwise_sum(we, n, 1); }

/* ------------------------------------------------------------------------------------------
 * resampling  (src/resample.jl)
 * ---------------------------------------------------------------------------------------- */
/* serial cumsum  resample.jl:19-22 / :40-43 */
static void cumsum_serial(const double* we, double* bins, int64_t N) {
  bins[0] = we[0];
  for (int64_t i = 1; i < N; ++i) bins[i] = bins[i - 1] + we[i];
}

/* ---- Julia Base's Float64 range `start:step:stop` (base/twiceprecision.jl) ----------------------------------
 * resample.jl:24 builds  s = r:(1/M):(bins[N]+r)  and compares s[i] with bins (:28).  Base is not under the
 * reference tree; the functions below restate the published algorithm of Julia 1.6 - 1.11 one to one
 * (rat, (:)(start,step,stop), floatrange, steprangelen_hp, TwicePrecision{Float64}((n,d)[,nb]), twiceprecision,
 * truncbits, nbitslen, add12, mul12, canonicalize2, TwicePrecision division, unsafe_getindex).  An independent
 * Python restatement lives in oracle/julia_range.py; tests/test_julia_range.py compares the two bit for bit.
 *   fallback path (start / stop not exact small rationals: the case for a 53-bit rand()):
 *       s[i] = fl(r + fl((i-1) * fl(1/M)))        (product rounded, then one add; no FMA)
 *   rational path (rand() == 0, dyadic test inputs, ~1e-3 of random draws at N = 10, ~1e-8 at N = 2^20):
 *       s[i] = double-double evaluation of (start_n + (i-1) step_n)/den  (can differ in the last bit)        */
typedef struct { double ref_hi, ref_lo, step_hi, step_lo; int64_t len, offset; int rational; } orc_range;

static void jl_rat(double x, int64_t* num, int64_t* den) {
  double y = x;
  int64_t a = 1, d = 1, b = 0, c = 0;
  const double m = 16777216.0;                     /* maxintfloat(narrow(Float64)) = maxintfloat(Float32) */
  while (fabs(y) <= m) {
    const int64_t f = (int64_t)y;                  /* trunc(Int, y) */
    y -= (double)f;
    const int64_t a2 = f * a + c, b2 = f * b + d;
    c = a; a = a2; d = b; b = b2;
    if (!((llabs(a) > llabs(b) ? llabs(a) : llabs(b)) <= (int64_t)m)) { *num = c; *den = d; return; }
    if ((double)a / (double)b == x) break;
    y = 1.0 / y;                                   /* inv(0.0) = Inf ends the loop */
  }
  *num = a; *den = b;
}
static int jl_isbetween(double a, double x, double b) { return (a <= x && x <= b) || (b <= x && x <= a); }
static void jl_canonicalize2(double big, double little, double* h, double* l) {
  *h = big + little;
  *l = (big - *h) + little;
}
static void jl_add12(double x, double y, double* h, double* l) {
  if (fabs(y) > fabs(x)) { const double t = x; x = y; y = t; }
  jl_canonicalize2(x, y, h, l);
}
static void jl_tp_div(double xhi, double xlo, double yhi, double ylo, double* h, double* l) {
  const double hi = xhi / yhi;
  const double uh = hi * yhi;
  const double ul = fma(hi, yhi, -uh);             /* mul12: exact product error */
  const double lo = ((((xhi - uh) - ul) + xlo) - hi * ylo) / yhi;
  jl_canonicalize2(hi, lo, h, l);
}
static double jl_truncbits(double x, int nb) {
  uint64_t u; memcpy(&u, &x, 8);
  u &= (nb >= 64) ? 0ull : (~0ull << nb);
  memcpy(&x, &u, 8);
  return x;
}
static int jl_nbitslen(int64_t len, int64_t offset) {
  if (len < 2) return 0;
  const int64_t a = offset - 1, b = len - offset;
  const int nb = (int)ceil(log2((double)(a > b ? a : b))) + 1;
  return nb < 27 ? nb : 27;                        /* min(cld(precision(Float64), 2), ...) */
}
static void jl_tp_ratio(int64_t n, int64_t d, int nb, double* h, double* l) {
  jl_tp_div((double)n, 0.0, (double)d, 0.0, h, l); /* |n|, |d| <= 2^53: exact conversions */
  if (nb >= 0) { const double h2 = jl_truncbits(*h, nb); *l = (*h - h2) + *l; *h = h2; }
}
static int64_t jl_gcd(int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return llabs(a); }

void orc_julia_range(double start, double step, double stop, orc_range* R) {
  int64_t step_n, step_d, start_n, start_d, stop_n, stop_d;
  jl_rat(step, &step_n, &step_d);
  if (step_d != 0 && (double)step_n / (double)step_d == step) {
    jl_rat(start, &start_n, &start_d);
    jl_rat(stop, &stop_n, &stop_d);
    if (start_d != 0 && stop_d != 0 && (double)start_n / (double)start_d == start &&
        (double)stop_n / (double)stop_d == stop) {
      const int64_t den = start_d / jl_gcd(start_d, step_d) * step_d;     /* lcm_unchecked; both <= 2^24 */
      const double m = 9007199254740992.0;                                /* maxintfloat(Float64) */
      if (den != 0 && fabs(start * (double)den) <= m && fabs(step * (double)den) <= m &&
          den % start_d == 0 && den % step_d == 0) {
        start_n = (int64_t)nearbyint(start * (double)den);
        step_n = (int64_t)nearbyint(step * (double)den);
        const __int128 q = ((__int128)den * stop_n) / stop_d;
        int64_t len = (int64_t)((q - start_n) / step_n) + 1;
        if (len < 0) len = 0;
        if (jl_isbetween(start, start + (double)(len - 1) * step, stop + step / 2) &&
            !jl_isbetween(start, start + (double)len * step, stop)) {
          R->rational = 1; R->len = len;
          if (len < 2 || step_n == 0) {
            R->offset = 1;
            jl_tp_ratio(start_n, den, -1, &R->ref_hi, &R->ref_lo);
            jl_tp_ratio(step_n, den, 0, &R->step_hi, &R->step_lo);
            return;
          }
          int64_t imin = (int64_t)nearbyint(-(double)start_n / (double)step_n + 1);
          if (imin < 1) imin = 1;
          if (imin > len) imin = len;
          const int64_t ref_n = start_n + (imin - 1) * step_n;
          R->offset = imin;
          jl_tp_ratio(ref_n, den, -1, &R->ref_hi, &R->ref_lo);
          jl_tp_ratio(step_n, den, jl_nbitslen(len, imin), &R->step_hi, &R->step_lo);
          return;
        }
      }
    }
  }
  const double lf = (stop - start) / step;
  int64_t len;
  if (lf < 0) len = 0;
  else if (lf == 0) len = 1;
  else {
    len = (int64_t)nearbyint(lf) + 1;
    const double stop2 = start + (double)(len - 1) * step;
    len -= (start < stop && stop < stop2) + (start > stop && stop > stop2);
  }
  R->rational = 0; R->len = len; R->offset = 1;
  R->ref_hi = start; R->ref_lo = 0.0; R->step_hi = step; R->step_lo = 0.0;
}
/* unsafe_getindex(r::StepRangeLen{Float64,TwicePrecision,TwicePrecision}, i), i 1-based, no bounds check
   (resample.jl:27-28 index under @inbounds) */
double orc_julia_range_getindex(const orc_range* R, int64_t i) {
  const double u = (double)(i - R->offset);
  const double shift_hi = u * R->step_hi, shift_lo = u * R->step_lo;
  double x_hi, x_lo;
  jl_add12(R->ref_hi, shift_hi, &x_hi, &x_lo);
  return x_hi + (x_lo + (shift_lo + R->ref_lo));
}

/* resample(ResampleSystematic, we, j, bins, M)  resample.jl:17-36, with rand() (:23) given as u01.
 * s = r:(1/M):(bins[N]+r) is Julia Base's Float64 range (above).  j is 1-based.  Entries for which no bin
 * satisfies s[i] < bins[b] keep their previous value. */
void orc_resample_systematic(const double* we, int64_t N, double u01, int64_t M, int64_t* j,
                             double* bins) {
  cumsum_serial(we, bins, N);
  const double r = u01 * bins[N - 1] / (double)N;
  orc_range s;
  orc_julia_range(r, 1.0 / (double)M, bins[N - 1] + r, &s);
  int64_t bo = 0;
  for (int64_t i = 0; i < M; ++i) {
    const double si = orc_julia_range_getindex(&s, i + 1);
    for (int64_t b = bo; b < N; ++b) {
      if (si < bins[b]) {
        j[i] = b + 1;
        bo = b;
        break;
      }
    }
  }
}

/* resample(ResampleStratified, ...)  resample.jl:38-61, the M rand() draws (:49) given as u01[M] */
void orc_resample_stratified(const double* we, int64_t N, const double* u01, int64_t M, int64_t* j,
                             double* bins) {
  cumsum_serial(we, bins, N);
  int64_t bo = 0;
  for (int64_t i = 0; i < M; ++i) {
    const double u = ((double)i + u01[i]) / (double)M * bins[N - 1];
    for (int64_t b = bo; b < N; ++b) {
      if (u < bins[b]) {
        j[i] = b + 1;
        bo = b;
        break;
      }
    }
  }
}

/* resample(ResampleResidual, ...)  resample.jl:63-117; the rand() at :106 is consumed from u01[]
 * in order (one per residual draw). Returns the number of uniforms consumed. */
int64_t orc_resample_residual(const double* we, int64_t N, const double* u01, int64_t M, int64_t* j,
                              double* bins) {
  double wsum = 0;
  for (int64_t i = 0; i < N; ++i) wsum += we[i];
  const double inv_wsum = 1 / wsum;
  int64_t num = 0;
  for (int64_t i = 0; i < N; ++i) {
    const double nw = we[i] * inv_wsum * (double)M;
    const int64_t cnt = (int64_t)floor(nw);
    bins[i] = nw - (double)cnt;
    for (int64_t k = 0; k < cnt; ++k) j[num++] = i + 1;
  }
  if (num == M) return 0;
  double rsum = 0;
  for (int64_t i = 0; i < N; ++i) rsum += bins[i];
  const double inv_rsum = 1 / rsum;
  for (int64_t i = 0; i < N; ++i) bins[i] *= inv_rsum;
  for (int64_t i = 1; i < N; ++i) bins[i] += bins[i - 1];
  int64_t used = 0;
  for (int64_t m = num; m < M; ++m) {
    const double u = u01[used++];
    for (int64_t i = 0; i < N; ++i) {
      if (u < bins[i]) { j[m] = i + 1; break; }
    }
  }
  return used;
}

/* ------------------------------------------------------------------------------------------
 * filter state  (src/PFtypes.jl:8-17 PFstate; :21-49, :162-177 the filter structs)
 * ---------------------------------------------------------------------------------------- */
typedef struct orc_pf {
  int64_t N;
  int nx, nu, ny;
  int filter, resampling, dynamics;
  double threshold, Ts;
  uint64_t seed, epoch;
  /* PFstate */
  double *x, *xprev;   /* AoS [N][nx] like Vector{SVector} */
  double *w, *we, *bins;
  int64_t* j;
  double maxw;
  int64_t t;
  /* model */
  double *A, *B, *C, *L1, *L2, *L0, *mu0;
  double c0_meas;      /* mvnormal_c0  utils.jl:254-257 */
  double dyn_params[8], t_switch, a1_factor, integ_Ts;
  int supersample;
  /* scratch */
  double* lambda;      /* APF: the reference aliases s.we; kept separately addressable here */
  int64_t resample_count;
  /* user closures (the reference's `dynamics(x,u,p,t)` / `measurement_likelihood(x,u,y,p,t)` are arbitrary Julia
     functions, PFtypes.jl:128,232): optional C callbacks replacing the descriptor model — test-side only, so that
     device-compiled user functions can be checked against the same function evaluated on the host */
  void (*user_dynamics)(const double* x, const double* u, double t, double* out, void* ctx);
  double (*user_loglik)(const double* x, const double* u, const double* y, double t, void* ctx);
  void* user_ctx;
  /* Float32-particle mode (llpf_config.particle_dtype == LLPF_PARTICLE_F32): see the f32 section below */
  int f32;
  float *Af, *Bf, *L1f, *L0f, *mu0f, *Gf;   /* row-major f32 copies; Gf = (float)(inv(chol(R2)) * C) */
  double* Wd;                               /* row-major lower inv(chol(R2)) */
  float c0f;
} orc_pf;

static double* dup_d(const double* p, size_t n) {
  double* q = (double*)malloc(sizeof(double) * (n ? n : 1));
  if (p && n) memcpy(q, p, sizeof(double) * n);
  return q;
}

void orc_destroy(orc_pf* f) {
  if (!f) return;
  free(f->x); free(f->xprev); free(f->w); free(f->we); free(f->bins); free(f->j);
  free(f->A); free(f->B); free(f->C); free(f->L1); free(f->L2); free(f->L0); free(f->mu0);
  free(f->lambda);
  free(f->Af); free(f->Bf); free(f->L1f); free(f->L0f); free(f->mu0f); free(f->Gf); free(f->Wd);
  free(f);
}

void orc_reset(orc_pf* f, uint64_t epoch);

/* ------------------------------------------------------------------------------------------
 * Float32 particles (BASELINE config 5: 64-state linear-Gaussian, test/test_large.jl:8-22 regime)
 *
 * The particle eltype of the reference comes from `initial_density` (PFtypes.jl:66,202); with Float32
 * densities and matrices `dynamics`, `rand!` and `logpdf` run in Float32 while w/we/bins stay Float64
 * (PFtypes.jl:68-69: `w = fill(log(1/N), N)`), so `w[i] += logpdf(...)` (PFtypes.jl:116) promotes a
 * Float32 log-likelihood.  At this size the reference's `A*x` is a BLAS sgemv and `logpdf` a PDMats
 * triangular solve, whose summation orders are unspecified; the order below is OURS (documented in
 * llpf_wide.cuh) and every operation is a correctly rounded f32 fmaf/add, so the CUDA path can match
 * it bit for bit.  Values are kept in the double arrays of orc_pf (exactly representable).
 *   x'_r   = fmaf-chain_c(A[r,c] x[c]) from 0 ; x'_r += (B u)_r ; x'_r = fmaf(L1[r,c], z[c], x'_r) for c <= r
 *   loglik = fmaf(-1/2, sum_a v_a^2 (fmaf chain), c0),  v_a = yt_a - (even-column chain + odd-column chain of G[a,:] x')
 *   yt = (float)(W y) with W = inv(chol(R2)) (f64), G = (float)(W C) (f64 product, rounded once)
 * ---------------------------------------------------------------------------------------- */
static void build_f32_model(orc_pf* f) {
  const int nx = f->nx, nu = f->nu, ny = f->ny;
  free(f->Af); free(f->Bf); free(f->L1f); free(f->L0f); free(f->mu0f); free(f->Gf); free(f->Wd);
  f->Af = (float*)calloc((size_t)nx * nx, sizeof(float));
  f->Bf = (float*)calloc((size_t)nx * (nu ? nu : 1), sizeof(float));
  f->L1f = (float*)calloc((size_t)nx * nx, sizeof(float));
  f->L0f = (float*)calloc((size_t)nx * nx, sizeof(float));
  f->mu0f = (float*)calloc((size_t)nx, sizeof(float));
  f->Gf = (float*)calloc((size_t)ny * nx, sizeof(float));
  f->Wd = (double*)calloc((size_t)ny * ny, sizeof(double));
  for (int r = 0; r < nx; ++r) {
    f->mu0f[r] = (float)f->mu0[r];
    for (int c = 0; c < nx; ++c) {
      f->Af[r * nx + c] = (float)CM(f->A, r, c, nx);
      f->L1f[r * nx + c] = (float)CM(f->L1, r, c, nx);
      f->L0f[r * nx + c] = (float)CM(f->L0, r, c, nx);
    }
    for (int c = 0; c < nu; ++c) f->Bf[r * nu + c] = (float)CM(f->B, r, c, nx);
  }
  /* W = inv(L2), column by column (forward substitution) */
  double* W = (double*)calloc((size_t)ny * ny, sizeof(double));   /* column-major */
  for (int c = 0; c < ny; ++c) {
    CM(W, c, c, ny) = 1.0 / CM(f->L2, c, c, ny);
    for (int r = c + 1; r < ny; ++r) {
      double acc = 0.0;
      for (int k = c; k < r; ++k) acc += CM(f->L2, r, k, ny) * CM(W, k, c, ny);
      CM(W, r, c, ny) = -acc / CM(f->L2, r, r, ny);
    }
  }
  for (int a = 0; a < ny; ++a) {
    for (int c = 0; c <= a; ++c) f->Wd[a * ny + c] = CM(W, a, c, ny);
    for (int c = 0; c < nx; ++c) {
      double acc = 0.0;
      for (int k = 0; k <= a; ++k) acc += CM(W, a, k, ny) * CM(f->C, k, c, ny);
      f->Gf[a * nx + c] = (float)acc;
    }
  }
  free(W);
  f->c0f = (float)f->c0_meas;
}

/* reset!(pf) filtering.jl:4-14, f32: x0_r = mu0_r + fmaf-chain_{c<=r}(L0[r,c] z[c]) */
static void reset_particles_f32(orc_pf* f) {
  const int nx = f->nx;
  double z[64];
  for (int64_t i = 0; i < f->N; ++i) {
    orc_normals(f->seed, f->epoch, STREAM_INIT, 0, (uint64_t)i, 64, z);   /* 16 counter blocks, as on the device */
    for (int r = 0; r < nx; ++r) {
      float acc = 0.f;
      for (int c = 0; c <= r; ++c) acc = fmaf(f->L0f[r * nx + c], (float)z[c], acc);
      acc = f->mu0f[r] + acc;
      f->xprev[(size_t)i * nx + r] = (double)acc;
      f->x[(size_t)i * nx + r] = (double)acc;
    }
  }
}

/* propagate_particles!  PFtypes.jl:122-139 / ext:83-93 / PFtypes.jl:242-289, f32 */
static void propagate_particles_f32(orc_pf* f, const double* u, int use_j, int with_noise) {
  const int nx = f->nx, nu = f->nu;
  float bu[64];
  for (int r = 0; r < nx; ++r) {
    float acc = 0.f;
    for (int c = 0; c < nu; ++c) acc = fmaf(f->Bf[r * nu + c], (float)u[c], acc);
    bu[r] = acc;
  }
  double z[64];
  for (int64_t i = 0; i < f->N; ++i) {
    const int64_t src = use_j ? f->j[i] - 1 : i;
    const double* xp = f->xprev + (size_t)src * nx;
    if (with_noise) orc_normals(f->seed, f->epoch, STREAM_DYN, (uint32_t)f->t, (uint64_t)i, 64, z);
    for (int r = 0; r < nx; ++r) {
      float acc = 0.f;
      for (int c = 0; c < nx; ++c) acc = fmaf(f->Af[r * nx + c], (float)xp[c], acc);
      acc = acc + bu[r];
      if (with_noise)
        for (int c = 0; c <= r; ++c) acc = fmaf(f->L1f[r * nx + c], (float)z[c], acc);
      f->x[(size_t)i * nx + r] = (double)acc;
    }
  }
}

/* measurement_equation!  PFtypes.jl:107-120 / :226-239, f32 log-likelihood promoted to f64 on `+=` */
static void measurement_equation_f32(const orc_pf* f, const double* y, double* w) {
  const int nx = f->nx, ny = f->ny;
  float yt[64];
  for (int a = 0; a < ny; ++a) {
    double acc = 0.0;
    for (int c = 0; c <= a; ++c) acc = fma(f->Wd[a * ny + c], y[c], acc);
    yt[a] = (float)acc;
  }
  for (int64_t i = 0; i < f->N; ++i) {
    const double* x = f->x + (size_t)i * nx;
    float q = 0.f;
    for (int a = 0; a < ny; ++a) {
      float de = 0.f, dd = 0.f;
      for (int c = 0; c < nx; c += 2) {
        de = fmaf(f->Gf[a * nx + c], (float)x[c], de);
        if (c + 1 < nx) dd = fmaf(f->Gf[a * nx + c + 1], (float)x[c + 1], dd);
      }
      const float v = yt[a] - (de + dd);
      q = fmaf(v, v, q);
    }
    w[i] += (double)fmaf(-0.5f, q, f->c0f);
  }
}

static int set_model(orc_pf* f, const llpf_model* m) {
  const int nx = m->nx, nu = m->nu, ny = m->ny;
  f->nx = nx; f->nu = nu; f->ny = ny; f->dynamics = m->dynamics;
  free(f->A); free(f->B); free(f->C); free(f->L1); free(f->L2); free(f->L0); free(f->mu0);
  f->A = dup_d(m->A, m->A ? (size_t)nx * nx : 0);
  f->B = dup_d(m->B, m->B ? (size_t)nx * nu : 0);
  f->C = dup_d(m->C, (size_t)ny * nx);
  f->mu0 = dup_d(m->mu0, nx);
  f->L1 = (double*)malloc(sizeof(double) * nx * nx);
  f->L2 = (double*)malloc(sizeof(double) * ny * ny);
  f->L0 = (double*)malloc(sizeof(double) * nx * nx);
  if (orc_cholesky_lower(m->R1, nx, f->L1)) return LLPF_ERR_NOT_POSDEF;
  if (orc_cholesky_lower(m->R2, ny, f->L2)) return LLPF_ERR_NOT_POSDEF;
  if (orc_cholesky_lower(m->Sigma0, nx, f->L0)) return LLPF_ERR_NOT_POSDEF;
  /* mvnormal_c0 = -(k*log2pi + logdet)/2   utils.jl:254-257 ; logdet = 2*sum(log(diag(L))) */
  double ld = 0;
  for (int i = 0; i < ny; ++i) ld += log(CM(f->L2, i, i, ny));
  ld *= 2;
  f->c0_meas = -((double)ny * log(2 * M_PI) + ld) / 2;
  memcpy(f->dyn_params, m->dyn_params, sizeof(f->dyn_params));
  f->t_switch = m->t_switch; f->a1_factor = m->a1_factor;
  f->integ_Ts = m->integ_Ts; f->supersample = m->supersample;
  if (f->f32) build_f32_model(f);
  return LLPF_OK;
}

/* constructors PFtypes.jl:65-75 / :200-210 / :38-49 (initial particles are drawn again by reset!) */
int orc_create(const llpf_config* cfg, const llpf_model* m, orc_pf** out) {
  if (!cfg || !m || !out || cfg->N <= 0) return LLPF_ERR_BAD_ARG;
  orc_pf* f = (orc_pf*)calloc(1, sizeof(orc_pf));
  f->N = cfg->N; f->filter = cfg->filter; f->resampling = cfg->resampling;
  f->threshold = cfg->resample_threshold; f->Ts = cfg->Ts; f->seed = cfg->seed;
  f->f32 = (cfg->particle_dtype == LLPF_PARTICLE_F32);
  if (f->f32 && (m->dynamics != LLPF_DYN_LINEAR || f->filter > LLPF_FILTER_ADVANCED)) { orc_destroy(f); return LLPF_ERR_UNSUPPORTED; }
  const int rc = set_model(f, m);
  if (rc) { orc_destroy(f); return rc; }
  const size_t N = (size_t)f->N;
  f->x = (double*)calloc(N * f->nx, sizeof(double));
  f->xprev = (double*)calloc(N * f->nx, sizeof(double));
  f->w = (double*)calloc(N, sizeof(double));
  f->we = (double*)calloc(N, sizeof(double));
  f->bins = (double*)calloc(N, sizeof(double));
  f->lambda = (double*)calloc(N, sizeof(double));
  f->j = (int64_t*)calloc(N, sizeof(int64_t));
  for (size_t i = 0; i < N; ++i) f->j[i] = (int64_t)i + 1; /* collect(1:N)  PFtypes.jl:70 */
  orc_reset(f, 0);
  f->t = 0; /* PFstate(...,Ref(0)) PFtypes.jl:70 ; reset! sets 1 */
  *out = f;
  return LLPF_OK;
}
int orc_set_model(orc_pf* f, const llpf_model* m) { return set_model(f, m); }
/* replace the dynamics mean and / or the measurement log-likelihood by caller-supplied functions (NULL keeps the
   descriptor's); Float64 particles only */
int orc_set_user_functions(orc_pf* f,
                           void (*dyn)(const double*, const double*, double, double*, void*),
                           double (*loglik)(const double*, const double*, const double*, double, void*), void* ctx) {
  if (f->f32) return LLPF_ERR_UNSUPPORTED;
  f->user_dynamics = dyn; f->user_loglik = loglik; f->user_ctx = ctx;
  return LLPF_OK;
}

/* rand(rng, d) = mu + L*z   utils.jl:260-262 ; Distributions MvNormal: mu + unwhiten(z) */
static void sample_mvn(const double* mu, const double* L, int n, const double* z, double* out) {
  for (int r = 0; r < n; ++r) {
    double acc = 0;
    for (int c = 0; c <= r; ++c) acc += CM(L, r, c, n) * z[c];
    out[r] = mu ? mu[r] + acc : acc;
  }
}

/* reset!(pf)  filtering.jl:4-14 */
void orc_reset(orc_pf* f, uint64_t epoch) {
  f->epoch = epoch;
  double z[64];
  if (f->f32) reset_particles_f32(f);
  else for (int64_t i = 0; i < f->N; ++i) {
    orc_normals(f->seed, epoch, STREAM_INIT, 0, (uint64_t)i, f->nx, z);
    sample_mvn(f->mu0, f->L0, f->nx, z, f->xprev + (size_t)i * f->nx);
    memcpy(f->x + (size_t)i * f->nx, f->xprev + (size_t)i * f->nx, sizeof(double) * f->nx);
  }
  const double lw = -log((double)f->N);
  const double ew = 1 / (double)f->N;
  for (int64_t i = 0; i < f->N; ++i) { f->w[i] = lw; f->we[i] = ew; }
  f->t = 1;
  f->resample_count = 0;
}

/* ------------------------------------------------------------------------------------------
 * models
 * ---------------------------------------------------------------------------------------- */
/* quadtank(h,u,p,t)  examples/example_quadtank.jl:91-106 (parametrised) with the optional
   `if t > 500; a1 *= 2` of the hard-coded variant :15-17 */
static void quadtank(const orc_pf* f, const double* h, const double* u, double t, double* xd) {
  const double* p = f->dyn_params;
  const double k1 = p[1], k2 = p[2], g = 9.81;
  const double A1 = p[3], A2 = p[3], A3 = p[3], A4 = p[3];
  double a1 = p[4];
  const double a2 = p[4], a3 = p[4], a4 = p[4];
  const double g1 = p[5], g2 = p[5];
  if (t > f->t_switch) a1 *= f->a1_factor;
#define SSQRT(v) sqrt(((v) > 0 ? (v) : 0.0) + 1e-3)
  const double tg = 2 * g;
  xd[0] = -a1 / A1 * SSQRT(tg * h[0]) + a3 / A1 * SSQRT(tg * h[2]) + g1 * k1 / A1 * u[0];
  xd[1] = -a2 / A2 * SSQRT(tg * h[1]) + a4 / A2 * SSQRT(tg * h[3]) + g2 * k2 / A2 * u[1];
  xd[2] = -a3 / A3 * SSQRT(tg * h[2]) + (1 - g2) * k2 / A3 * u[1];
  xd[3] = -a4 / A4 * SSQRT(tg * h[3]) + (1 - g1) * k1 / A4 * u[0];
#undef SSQRT
}

/* rk4(f, Ts0; supersample)  utils.jl:220-237, generic over the right-hand side */
typedef void (*rhs_fn)(const void* ctx, const double* x, const double* u, double t, double* xd);
static void rk4_generic(rhs_fn rhs, const void* ctx, int n, double Ts0, int supersample,
                        const double* x0, const double* u, double t, double* xout) {
  const double Ts = Ts0 / (double)supersample;
  double x[64], f1[64], f2[64], f3[64], f4[64], tmp[64];
  memcpy(x, x0, sizeof(double) * n);
  for (int s = 0; s < supersample; ++s) {
    rhs(ctx, x, u, t, f1);
    for (int i = 0; i < n; ++i) tmp[i] = x[i] + Ts / 2 * f1[i];
    rhs(ctx, tmp, u, t + Ts / 2, f2);
    for (int i = 0; i < n; ++i) tmp[i] = x[i] + Ts / 2 * f2[i];
    rhs(ctx, tmp, u, t + Ts / 2, f3);
    for (int i = 0; i < n; ++i) tmp[i] = x[i] + Ts * f3[i];
    rhs(ctx, tmp, u, t + Ts, f4);
    for (int i = 0; i < n; ++i) x[i] += Ts / 6 * (f1[i] + 2 * f2[i] + 2 * f3[i] + f4[i]);
    t += Ts;
  }
  memcpy(xout, x, sizeof(double) * n);
}
static void quadtank_rhs(const void* ctx, const double* x, const double* u, double t, double* xd) {
  quadtank((const orc_pf*)ctx, x, u, t, xd);
}
static void rk4_quadtank(const orc_pf* f, const double* x0, const double* u, double t, double* xout) {
  rk4_generic(quadtank_rhs, f, 4, f->integ_Ts, f->supersample, x0, u, t, xout);
}
/* test hook for the reference's rk4 known answer (test/runtests.jl:182-188): xdot = c (constant) */
static void const_rhs(const void* ctx, const double* x, const double* u, double t, double* xd) {
  (void)x; (void)u; (void)t;
  xd[0] = *(const double*)ctx;
}
double orc_rk4_constant_rhs(double c, double x0, double Ts, int supersample) {
  double out;
  rk4_generic(const_rhs, &c, 1, Ts, supersample, &x0, NULL, 1.0, &out);
  return out;
}

/* dynamics(x,u,p,t) without noise: LG `A*x .+ B*u` (example_lineargaussian.jl:27) or rk4(quadtank) */
static void dynamics_mean(const orc_pf* f, const double* x, const double* u, double t, double* out) {
  if (f->user_dynamics) { f->user_dynamics(x, u, t, out, f->user_ctx); return; }
  if (f->dynamics == LLPF_DYN_QUADTANK_RK4) {
    rk4_quadtank(f, x, u, t, out);
    return;
  }
  double ax[64], bu[64];
  matvec(f->A, f->nx, f->nx, x, ax);
  if (f->nu > 0) {
    matvec(f->B, f->nx, f->nu, u, bu);
    for (int r = 0; r < f->nx; ++r) out[r] = ax[r] + bu[r];
  } else {
    for (int r = 0; r < f->nx; ++r) out[r] = ax[r];
  }
}

/* extended_logpdf(d, r) = mvnormal_c0(d) - invquad(Sigma, r)/2   utils.jl:252-257 ;
   invquad through the Cholesky factor (PDMats): |L \ r|^2 */
static double meas_logpdf(const orc_pf* f, const double* r) {
  double v[64];
  const int n = f->ny;
  double q = 0;
  for (int i = 0; i < n; ++i) {
    double acc = r[i];
    for (int k = 0; k < i; ++k) acc -= CM(f->L2, i, k, n) * v[k];
    v[i] = acc / CM(f->L2, i, i, n);
    q += v[i] * v[i];
  }
  return f->c0_meas - q / 2;
}

static int any_nan(const double* y, int n) {
  for (int i = 0; i < n; ++i) if (isnan(y[i])) return 1;
  return 0;
}

/* measurement_equation!(pf,u,y,p,t,w)  PFtypes.jl:107-120 (PF): w[i] += logpdf(dg, y - g(x[i]))
   and PFtypes.jl:226-239 (Advanced): w[i] += g(x[i],u,y,p,t) with the Gaussian likelihood of
   example_lineargaussian.jl:238-240.  NaN in y == `missing` (:109, :227). */
static void measurement_equation(const orc_pf* f, const double* u, const double* y, double t, double* w) {
  (void)u; (void)t;
  if (any_nan(y, f->ny)) return;
  if (f->f32) { measurement_equation_f32(f, y, w); return; }
  if (f->user_loglik) {      /* w[i] += g(x[i],u,y,p,t)   PFtypes.jl:232 */
    for (int64_t i = 0; i < f->N; ++i) w[i] += f->user_loglik(f->x + (size_t)i * f->nx, u, y, t, f->user_ctx);
    return;
  }
  double g[64], r[64];
  for (int64_t i = 0; i < f->N; ++i) {
    matvec(f->C, f->ny, f->nx, f->x + (size_t)i * f->nx, g);
    for (int k = 0; k < f->ny; ++k) r[k] = y[k] - g[k];
    w[i] += meas_logpdf(f, r);
  }
}

/* dynamics-noise draw for (step index, particle): rand!(rng, d, noise) = L*z  PFtypes.jl:135,153 */
static void dyn_noise(const orc_pf* f, int64_t i, double* noise) {
  double z[64];
  orc_normals(f->seed, f->epoch, STREAM_DYN, (uint32_t)f->t, (uint64_t)i, f->nx, z);
  sample_mvn(NULL, f->L1, f->nx, z, noise);
}

/* propagate_particles!(pf,u,j,p,t,d)            PFtypes.jl:122-139  (PF, resampled)
   propagate_particles!(pf,u,p,t,d::Sampleable)  ext/LowLevelParticleFiltersDistributionsExt.jl:83-93 (PF, no resample)
   propagate_particles!(pf::Advanced,u,j,p,t,noise) PFtypes.jl:242-259, (pf,u,p,t,noise::Bool) :276-289
   propagate_particles!(pf,u,p,t,::Nothing)      PFtypes.jl:261-274  (noise-free)
   use_j: gather through j (1-based); with_noise: add L*z                                        */
static void propagate_particles(orc_pf* f, const double* u, int use_j, double t, int with_noise) {
  double fx[64], nz[64];
  if (f->f32) { propagate_particles_f32(f, u, use_j, with_noise); return; }
  for (int64_t i = 0; i < f->N; ++i) {
    const int64_t src = use_j ? f->j[i] - 1 : i;
    dynamics_mean(f, f->xprev + (size_t)src * f->nx, u, t, fx);
    if (with_noise) {
      dyn_noise(f, i, nz);
      for (int k = 0; k < f->nx; ++k) f->x[(size_t)i * f->nx + k] = fx[k] + nz[k];
    } else {
      for (int k = 0; k < f->nx; ++k) f->x[(size_t)i * f->nx + k] = fx[k];
    }
  }
}

/* add_noise!(pf)  PFtypes.jl:146-157 */
static void add_noise(orc_pf* f) {
  double nz[64];
  for (int64_t i = 0; i < f->N; ++i) {
    dyn_noise(f, i, nz);
    for (int k = 0; k < f->nx; ++k) f->x[(size_t)i * f->nx + k] += nz[k];
  }
}

/* reset_weights!(s)  utils.jl:73-78 */
static void reset_weights(orc_pf* f) {
  const double lw = log(1 / (double)f->N);
  const double ew = 1 / (double)f->N;
  for (int64_t i = 0; i < f->N; ++i) { f->w[i] = lw; f->we[i] = ew; }
  f->maxw = 0;
}

/* shouldresample(pf)  resample.jl:5-10 */
int orc_shouldresample(const orc_pf* f) {
  if (f->threshold == 1) return 1;
  const double th = (double)f->N * f->threshold;
  const double ne = orc_effective_particles(f->we, f->N);
  return ne < th;
}

/* resample(strategy, we, j, bins)  resample.jl:12-15 dispatch; rand() from the counter streams */
static void resample_dispatch(orc_pf* f, const double* we) {
  if (f->resampling == LLPF_RESAMPLE_RESIDUAL) {
    /* draw k (0-based, in the order of the rand() calls at resample.jl:106) = counter k of STREAM_RESID */
    double* u = (double*)malloc(sizeof(double) * f->N);
    for (int64_t i = 0; i < f->N; ++i)
      u[i] = orc_uniform53(f->seed, f->epoch, STREAM_RESID, (uint32_t)f->t, (uint64_t)i);
    orc_resample_residual(we, f->N, u, f->N, f->j, f->bins);
    free(u);
  } else if (f->resampling == LLPF_RESAMPLE_STRATIFIED) {
    double* u = (double*)malloc(sizeof(double) * f->N);
    for (int64_t i = 0; i < f->N; ++i)
      u[i] = orc_uniform53(f->seed, f->epoch, STREAM_STRAT, (uint32_t)f->t, (uint64_t)i);
    orc_resample_stratified(we, f->N, u, f->N, f->j, f->bins);
    free(u);
  } else {
    const double u01 = orc_uniform53(f->seed, f->epoch, STREAM_RESAMPLE, (uint32_t)f->t, 0);
    orc_resample_systematic(we, f->N, u01, f->N, f->j, f->bins);
  }
  f->resample_count++;
}

/* predict!(pf,u,p,t)  filtering.jl:140-153 */
void orc_predict(orc_pf* f, const double* u, double t) {
  if (orc_shouldresample(f)) {
    resample_dispatch(f, f->we);
    propagate_particles(f, u, 1, t, 1);
    reset_weights(f);
  } else {
    for (int64_t i = 0; i < f->N; ++i) f->j[i] = i + 1;
    propagate_particles(f, u, 0, t, 1);
  }
  memcpy(f->xprev, f->x, sizeof(double) * (size_t)f->N * f->nx);
  f->t += 1;
}

/* correct!(pf,u,y,p,t)  filtering.jl:164-168 ; APF :170-174 (logsumexp! only) */
double orc_correct(orc_pf* f, const double* u, const double* y, double t) {
  if (f->filter != LLPF_FILTER_AUX && f->filter != LLPF_FILTER_AUX_ADVANCED)
    measurement_equation(f, u, y, t, f->w);
  return orc_logsumexp(f->w, f->we, f->N, &f->maxw);
}

/* permute_with_buffer!(x, buf, j)  utils.jl:81-86 */
static void permute_with_buffer(orc_pf* f) {
  for (int64_t i = 0; i < f->N; ++i)
    memcpy(f->xprev + (size_t)i * f->nx, f->x + (size_t)(f->j[i] - 1) * f->nx, sizeof(double) * f->nx);
  memcpy(f->x, f->xprev, sizeof(double) * (size_t)f->N * f->nx);
}

/* predict!(pf::AuxiliaryParticleFilter,u,y1,p,t)  filtering.jl:195-217 ;
   AuxiliaryParticleFilter{<:AdvancedParticleFilter} :219-234 */
void orc_predict_aux(orc_pf* f, const double* u, const double* y1, double t) {
  propagate_particles(f, u, 0, t, 0);                 /* :199 / :221 propagate without noise */
  double* lam = f->we;                                /* :200 λ = s.we (alias) */
  for (int64_t i = 0; i < f->N; ++i) lam[i] = 0;      /* :201 */
  measurement_equation(f, u, y1, t, lam);             /* :202 */
  for (int64_t i = 0; i < f->N; ++i) f->w[i] += lam[i]; /* :203 */
  orc_expnormalize1(f->w, f->N);                      /* :204 w used as buffer */
  resample_dispatch(f, f->w);                         /* :205 */
  if (f->filter == LLPF_FILTER_AUX_ADVANCED) {
    reset_weights(f);                                 /* :228 (overwrites λ: we aliases it) */
    propagate_particles(f, u, 1, t, 1);               /* :230 with noise and permutation */
  } else {
    permute_with_buffer(f);                           /* :207 */
    add_noise(f);                                     /* :208 */
    const double lN = log((double)f->N);
    for (int64_t i = 0; i < f->N; ++i) f->w[i] = lam[i] - lN; /* :210-213 unresampled λ[i] */
  }
  f->t += 1;                                          /* :215 */
  memcpy(f->xprev, f->x, sizeof(double) * (size_t)f->N * f->nx); /* :216 */
}

/* update!(f,u,y,p,t) filtering.jl:181-185 ; update!(pfa,u,y,y1,p,t) :187-191 */
double orc_update(orc_pf* f, const double* u, const double* y, const double* y1, double t) {
  const double ll = orc_correct(f, u, y, t);
  if (f->filter == LLPF_FILTER_AUX || f->filter == LLPF_FILTER_AUX_ADVANCED)
    orc_predict_aux(f, u, y1, t);
  else
    orc_predict(f, u, t);
  return ll;
}

/* weighted_mean(x,we)  filtering.jl:541-548 */
void orc_weighted_mean(const orc_pf* f, double* xh) {
  for (int k = 0; k < f->nx; ++k) xh[k] = 0;
  for (int64_t i = 0; i < f->N; ++i)
    for (int k = 0; k < f->nx; ++k) xh[k] += f->x[(size_t)i * f->nx + k] * f->we[i];
}

/* forward_trajectory(pf,u,y,p)  filtering.jl:343-365 ; APF :367-384.
   u: [T][nu], y: [T][ny]; optional outputs (may be NULL): ll_steps[T], ess[T], resampled[T],
   xhat[T][nx], x_hist[T][N][nx], w_hist[T][N], we_hist[T][N].  Returns ll. */
double orc_forward_trajectory(orc_pf* f, int64_t T, const double* u, const double* y, uint64_t epoch,
                              double* ll_steps, double* ess, int32_t* resampled, double* xhat,
                              double* x_hist, double* w_hist, double* we_hist) {
  orc_reset(f, epoch);
  const size_t N = (size_t)f->N;
  const int aux = (f->filter == LLPF_FILTER_AUX || f->filter == LLPF_FILTER_AUX_ADVANCED);
  double ll = 0.;
  for (int64_t t = 0; t < T; ++t) {
    const double ti = (double)t * f->Ts;                 /* ti = (t-1)*Ts, t 1-based  :352 */
    const double* ut = u + (size_t)t * f->nu;
    const double* yt = y + (size_t)t * f->ny;
    const double lli = orc_correct(f, ut, yt, ti);
    ll += lli;
    if (ll_steps) ll_steps[t] = lli;
    if (ess) ess[t] = orc_effective_particles(f->we, f->N);
    if (xhat) orc_weighted_mean(f, xhat + (size_t)t * f->nx);
    if (x_hist) memcpy(x_hist + (size_t)t * N * f->nx, f->x, sizeof(double) * N * f->nx);
    if (w_hist) memcpy(w_hist + (size_t)t * N, f->w, sizeof(double) * N);
    if (we_hist) memcpy(we_hist + (size_t)t * N, f->we, sizeof(double) * N);
    const int64_t rc0 = f->resample_count;
    if (aux) {
      if (t < T - 1) orc_predict_aux(f, ut, y + (size_t)(t + 1) * f->ny, ti); /* :382 */
    } else {
      orc_predict(f, ut, ti);
    }
    if (resampled) resampled[t] = (int32_t)(f->resample_count - rc0);
  }
  return ll;
}

/* loglik(f,u,y,p)  smoothing.jl:227-230: reset!, then sum of f(u_t,y_t,p)[1] with the default
   t = index(pf)*Ts (filtering.jl:181,238).  APF: smoothing.jl:232-236 (explicit t=(t-1)*Ts, inner-PF tail). */
double orc_loglik(orc_pf* f, int64_t T, const double* u, const double* y, uint64_t epoch,
                  double* ll_steps, double* ess, int32_t* resampled) {
  orc_reset(f, epoch);
  const int aux = (f->filter == LLPF_FILTER_AUX || f->filter == LLPF_FILTER_AUX_ADVANCED);
  double ll = 0;
  for (int64_t t = 0; t < T; ++t) {
    const double* ut = u + (size_t)t * f->nu;
    const double* yt = y + (size_t)t * f->ny;
    const int64_t rc0 = f->resample_count;
    double lli;
    if (!aux) {
      const double ti = (double)f->t * f->Ts;
      lli = orc_correct(f, ut, yt, ti);
      if (ess) ess[t] = orc_effective_particles(f->we, f->N);
      orc_predict(f, ut, ti);
    } else if (t < T - 1) {
      const double ti = (double)t * f->Ts;
      lli = orc_correct(f, ut, yt, ti);
      if (ess) ess[t] = orc_effective_particles(f->we, f->N);
      orc_predict_aux(f, ut, y + (size_t)(t + 1) * f->ny, ti);
    } else {
      /* pf.pf(u[end], y[end], p, (length(u)-1)*Ts): the INNER filter's update!  smoothing.jl:235 */
      const double ti = (double)(T - 1) * f->Ts;
      const int keep = f->filter;
      f->filter = (keep == LLPF_FILTER_AUX) ? LLPF_FILTER_PF : LLPF_FILTER_ADVANCED;
      lli = orc_correct(f, ut, yt, ti);
      if (ess) ess[t] = orc_effective_particles(f->we, f->N);
      orc_predict(f, ut, ti);
      f->filter = keep;
    }
    ll += lli;
    if (ll_steps) ll_steps[t] = lli;
    if (resampled) resampled[t] = (int32_t)(f->resample_count - rc0);
  }
  return ll;
}

/* accessors */
/* ------------------------------------------------------------------------------------------
 * particle smoother: forward filtering, backward simulation   src/smoothing.jl:104-143
 * ---------------------------------------------------------------------------------------- */
/* logpdf(df, x, xp, t) = logpdf(df, x - xp)  ext/...DistributionsExt.jl:15 ; the Gaussian as in utils.jl:252-257 */
static double dyn_logpdf(const orc_pf* f, double c0_dyn, const double* r) {
  double v[64];
  const int n = f->nx;
  double q = 0;
  for (int i = 0; i < n; ++i) {
    double acc = r[i];
    for (int k = 0; k < i; ++k) acc -= CM(f->L1, i, k, n) * v[k];
    v[i] = acc / CM(f->L1, i, i, n);
    q += v[i] * v[i];
  }
  return c0_dyn - q / 2;
}

/* draw_one_categorical(pf, w)  resample.jl:128-152 : w are LOG-weights (normalised in place by logsumexp!),
   bins <- cumsum(exp weights), s = rand()*bins[end], two-sided linear search around the midpoint. 1-based result. */
static int64_t draw_one_categorical(double* w, double* bins, int64_t N, double u01) {
  orc_logsumexp(w, bins, N, NULL);
  for (int64_t i = 1; i < N; ++i) bins[i] += bins[i - 1];
  const double s = u01 * bins[N - 1];
  const int64_t midpoint = N / 2;                        /* 1-based index length(bins)÷2 */
  if (midpoint >= 1 && s < bins[midpoint - 1]) {
    for (int64_t b = 1; b <= midpoint; ++b)
      if (s <= bins[b - 1]) return b;
  } else {
    for (int64_t b = (midpoint >= 1 ? midpoint : 1); b <= N; ++b)
      if (s <= bins[b - 1]) return b;
  }
  return N;
}

/* smooth(pf, xf, wf, wef, ll, M, u, y, p)  smoothing.jl:116-143.
   xf [T][N][nx], wf / wef [T][N] (the ParticleFilteringSolution fields), xb out [T][M][nx] (M x T Matrix{SVector}).
   rand() draws: the initial resample uses counter step 0 of STREAM_SMOOTH with the strategy's usual indexing
   (systematic: index 0; stratified: slot i; residual: draw k); draw_one_categorical at (t, m), t 1-based: step t,
   index m (0-based).  Returns 0, or LLPF_ERR_BAD_ARG when M > N (the reference asserts, :122). */
int orc_smooth(orc_pf* f, int64_t T, int64_t M, const double* u, const double* xf, const double* wf,
               const double* wef, uint64_t epoch, double* xb) {
  const int64_t N = f->N;
  const int nx = f->nx;
  if (M > N || M < 1 || T < 1) return LLPF_ERR_BAD_ARG;
  double ld = 0;
  for (int i = 0; i < nx; ++i) ld += log(CM(f->L1, i, i, nx));
  ld *= 2;
  const double c0_dyn = -((double)nx * log(2 * M_PI) + ld) / 2;
  /* j = resample(pf.resampling_strategy, wef[:,T], M)  :124 -> resample.jl:14 (fresh j = zeros(Int,M), bins = zeros(N)) */
  int64_t* j = (int64_t*)calloc((size_t)M, sizeof(int64_t));
  double* bins = (double*)calloc((size_t)N, sizeof(double));
  const double* weT = wef + (size_t)(T - 1) * N;
  if (f->resampling == LLPF_RESAMPLE_SYSTEMATIC) {
    orc_resample_systematic(weT, N, orc_uniform53(f->seed, epoch, STREAM_SMOOTH, 0, 0), M, j, bins);
  } else {
    double* us = (double*)malloc(sizeof(double) * (size_t)M);
    for (int64_t i = 0; i < M; ++i) us[i] = orc_uniform53(f->seed, epoch, STREAM_SMOOTH, 0, (uint64_t)i);
    if (f->resampling == LLPF_RESAMPLE_STRATIFIED) orc_resample_stratified(weT, N, us, M, j, bins);
    else orc_resample_residual(weT, N, us, M, j, bins);
    free(us);
  }
  for (int64_t i = 0; i < M; ++i) {
    /* a slot the resampler left untouched holds 0 (zeros(Int,M)): xf[0,T] is a BoundsError in the reference;
       the restatement takes the last particle instead (cannot happen for weights that sum to one) */
    const int64_t a = (j[i] >= 1 && j[i] <= N) ? j[i] : N;
    memcpy(xb + ((size_t)(T - 1) * M + i) * nx, xf + ((size_t)(T - 1) * N + (a - 1)) * nx, sizeof(double) * nx);
  }
  double* wb = (double*)malloc(sizeof(double) * (size_t)N);
  double* fx = (double*)malloc(sizeof(double) * (size_t)N * nx);
  double r[64];
  for (int64_t t = T - 1; t >= 1; --t) {                 /* t is the reference's 1-based index, :130 */
    const double ti = (double)(t - 1) * f->Ts;           /* :131 */
    const double* ut = u + (size_t)(t - 1) * f->nu;
    const double* xft = xf + (size_t)(t - 1) * N * nx;
    const double* wft = wf + (size_t)(t - 1) * N;
    /* f(xf[n,t],u[t],p,ti) does not depend on m: evaluated once per n (same values as the reference's M evaluations) */
    for (int64_t n = 0; n < N; ++n) dynamics_mean(f, xft + (size_t)n * nx, ut, ti, fx + (size_t)n * nx);
    for (int64_t m = 0; m < M; ++m) {
      const double* xbn = xb + ((size_t)t * M + m) * nx; /* xb[m,t+1] */
      for (int64_t n = 0; n < N; ++n) {
        for (int k = 0; k < nx; ++k) r[k] = xbn[k] - fx[(size_t)n * nx + k];
        wb[n] = wft[n] + dyn_logpdf(f, c0_dyn, r);       /* :135 */
      }
      const int64_t i = draw_one_categorical(wb, bins, N, orc_uniform53(f->seed, epoch, STREAM_SMOOTH, (uint32_t)t, (uint64_t)m));
      memcpy(xb + ((size_t)(t - 1) * M + m) * nx, xft + (size_t)(i - 1) * nx, sizeof(double) * nx);   /* :139 */
    }
  }
  free(wb); free(fx); free(j); free(bins);
  return LLPF_OK;
}

int64_t orc_num_particles(const orc_pf* f) { return f->N; }
int64_t orc_index(const orc_pf* f) { return f->t; }
double* orc_particles(orc_pf* f) { return f->x; }
double* orc_xprev(orc_pf* f) { return f->xprev; }
double* orc_weights(orc_pf* f) { return f->w; }
double* orc_expweights(orc_pf* f) { return f->we; }
double* orc_bins(orc_pf* f) { return f->bins; }
int64_t* orc_ancestors(orc_pf* f) { return f->j; }
void orc_set_state(orc_pf* f, const double* x, const double* w, int64_t t) {
  const size_t n = (size_t)f->N * f->nx;
  memcpy(f->x, x, sizeof(double) * n);
  memcpy(f->xprev, x, sizeof(double) * n);
  memcpy(f->w, w, sizeof(double) * (size_t)f->N);
  for (int64_t i = 0; i < f->N; ++i) f->we[i] = exp(w[i]);
  f->t = t;
}

/* ------------------------------------------------------------------------------------------
 * data generation: simulate(f,u,p)  filtering.jl:462-477 with sample_state / sample_measurement
 * PFtypes.jl:302-306 (x1 = mean(d0), sample_initial=false).  Noise from STREAM_SIM.
 * ---------------------------------------------------------------------------------------- */
void orc_simulate(orc_pf* f, int64_t T, const double* u, uint64_t sim_seed, double* xs, double* ys) {
  double z[64], nz[64], g[64], fx[64];
  double* x = xs;
  memcpy(x, f->mu0, sizeof(double) * f->nx);
  for (int64_t t = 0; t < T; ++t) {
    const double ti = (double)t * f->Ts;
    double* xt = xs + (size_t)t * f->nx;
    /* y[t] = measurement(x[t]) + rand(dg) */
    matvec(f->C, f->ny, f->nx, xt, g);
    orc_normals(sim_seed, 0, STREAM_SIM, (uint32_t)t, 1, f->ny, z);
    sample_mvn(NULL, f->L2, f->ny, z, nz);
    for (int k = 0; k < f->ny; ++k) ys[(size_t)t * f->ny + k] = g[k] + nz[k];
    if (t < T - 1) {
      dynamics_mean(f, xt, u + (size_t)t * f->nu, ti, fx);
      orc_normals(sim_seed, 0, STREAM_SIM, (uint32_t)t, 0, f->nx, z);
      sample_mvn(NULL, f->L1, f->nx, z, nz);
      for (int k = 0; k < f->nx; ++k) xs[(size_t)(t + 1) * f->nx + k] = fx[k] + nz[k];
    }
  }
}

/* one noise-free dynamics evaluation (used to pin rk4 against test/runtests.jl:182-188) */
void orc_dynamics(orc_pf* f, const double* x, const double* u, double t, double* out) {
  dynamics_mean(f, x, u, t, out);
}

/* ------------------------------------------------------------------------------------------
 * closed-form Kalman filter log-likelihood for the LG model: the RNG-independent ground truth
 * the reference itself tests the PF against (test/runtests.jl:412-449).
 * reset! kalman.jl:159-164 (x=mu0, R=Sigma0); correct! filtering.jl:100-128; predict! :52-74;
 * loglik smoothing.jl:227-230.  Dense, small, column-major.
 * ---------------------------------------------------------------------------------------- */
double orc_kalman_loglik(const llpf_model* m, int64_t T, const double* u, const double* y) {
  const int nx = m->nx, nu = m->nu, ny = m->ny;
  double *x = dup_d(m->mu0, nx), *R = dup_d(m->Sigma0, (size_t)nx * nx);
  double *tmp = (double*)malloc(sizeof(double) * nx * nx), *tmp2 = (double*)malloc(sizeof(double) * nx * nx);
  double *S = (double*)malloc(sizeof(double) * ny * ny), *Ls = (double*)malloc(sizeof(double) * ny * ny);
  double *RCt = (double*)malloc(sizeof(double) * nx * ny), *K = (double*)malloc(sizeof(double) * nx * ny);
  double e[64], v[64], xn[64];
  double ll = 0;
  for (int64_t t = 0; t < T; ++t) {
    const double* ut = u + (size_t)t * nu;
    const double* yt = y + (size_t)t * ny;
    /* e = y - C x */
    for (int i = 0; i < ny; ++i) {
      double acc = 0;
      for (int k = 0; k < nx; ++k) acc += CM(m->C, i, k, ny) * x[k];
      e[i] = yt[i] - acc;
    }
    /* RCt = R C' ; S = C R C' + R2 */
    for (int i = 0; i < nx; ++i)
      for (int jn = 0; jn < ny; ++jn) {
        double acc = 0;
        for (int k = 0; k < nx; ++k) acc += CM(R, i, k, nx) * CM(m->C, jn, k, ny);
        CM(RCt, i, jn, nx) = acc;
      }
    for (int i = 0; i < ny; ++i)
      for (int jn = 0; jn < ny; ++jn) {
        double acc = 0;
        for (int k = 0; k < nx; ++k) acc += CM(m->C, i, k, ny) * CM(RCt, k, jn, nx);
        CM(S, i, jn, ny) = acc;
      }
    for (int i = 0; i < ny; ++i)
      for (int jn = i + 1; jn < ny; ++jn) {
        const double a = 0.5 * (CM(S, i, jn, ny) + CM(S, jn, i, ny));
        CM(S, i, jn, ny) = a; CM(S, jn, i, ny) = a;
      }
    for (int i = 0; i < ny * ny; ++i) S[i] += m->R2[i];
    if (orc_cholesky_lower(S, ny, Ls)) { ll = NAN; break; }
    /* ll += logpdf(N(0,S), e) */
    double q = 0, ld = 0;
    for (int i = 0; i < ny; ++i) {
      double acc = e[i];
      for (int k = 0; k < i; ++k) acc -= CM(Ls, i, k, ny) * v[k];
      v[i] = acc / CM(Ls, i, i, ny);
      q += v[i] * v[i];
      ld += log(CM(Ls, i, i, ny));
    }
    ll += -((double)ny * log(2 * M_PI) + 2 * ld) / 2 - q / 2;
    /* K = RCt / S  (solve K S = RCt via the Cholesky factor) */
    for (int i = 0; i < nx; ++i) {
      double z1[64], z2[64];
      for (int c = 0; c < ny; ++c) { /* forward: z1 L' = row  -> L z1' = row' */
        double acc = CM(RCt, i, c, nx);
        for (int k = 0; k < c; ++k) acc -= CM(Ls, c, k, ny) * z1[k];
        z1[c] = acc / CM(Ls, c, c, ny);
      }
      for (int c = ny - 1; c >= 0; --c) {
        double acc = z1[c];
        for (int k = c + 1; k < ny; ++k) acc -= CM(Ls, k, c, ny) * z2[k];
        z2[c] = acc / CM(Ls, c, c, ny);
      }
      for (int c = 0; c < ny; ++c) CM(K, i, c, nx) = z2[c];
    }
    for (int i = 0; i < nx; ++i) {
      double acc = 0;
      for (int c = 0; c < ny; ++c) acc += CM(K, i, c, nx) * e[c];
      x[i] += acc;
    }
    /* R = symmetrize((I - K C) R) */
    for (int i = 0; i < nx; ++i)
      for (int jn = 0; jn < nx; ++jn) {
        double acc = (i == jn) ? 1.0 : 0.0;
        for (int c = 0; c < ny; ++c) acc -= CM(K, i, c, nx) * CM(m->C, c, jn, ny);
        CM(tmp, i, jn, nx) = acc;
      }
    for (int i = 0; i < nx; ++i)
      for (int jn = 0; jn < nx; ++jn) {
        double acc = 0;
        for (int k = 0; k < nx; ++k) acc += CM(tmp, i, k, nx) * CM(R, k, jn, nx);
        CM(tmp2, i, jn, nx) = acc;
      }
    for (int i = 0; i < nx; ++i)
      for (int jn = 0; jn < nx; ++jn) CM(R, i, jn, nx) = 0.5 * (CM(tmp2, i, jn, nx) + CM(tmp2, jn, i, nx));
    /* predict: x = A x + B u ; R = A R A' + R1 */
    for (int i = 0; i < nx; ++i) {
      double acc = 0;
      for (int k = 0; k < nx; ++k) acc += CM(m->A, i, k, nx) * x[k];
      for (int k = 0; k < nu; ++k) acc += CM(m->B, i, k, nx) * ut[k];
      xn[i] = acc;
    }
    memcpy(x, xn, sizeof(double) * nx);
    for (int i = 0; i < nx; ++i)
      for (int jn = 0; jn < nx; ++jn) {
        double acc = 0;
        for (int k = 0; k < nx; ++k) acc += CM(m->A, i, k, nx) * CM(R, k, jn, nx);
        CM(tmp, i, jn, nx) = acc;
      }
    for (int i = 0; i < nx; ++i)
      for (int jn = 0; jn < nx; ++jn) {
        double acc = 0;
        for (int k = 0; k < nx; ++k) acc += CM(tmp, i, k, nx) * CM(m->A, jn, k, nx);
        CM(tmp2, i, jn, nx) = acc;
      }
    for (int i = 0; i < nx; ++i)
      for (int jn = 0; jn < nx; ++jn)
        CM(R, i, jn, nx) = 0.5 * (CM(tmp2, i, jn, nx) + CM(tmp2, jn, i, nx)) + CM(m->R1, i, jn, nx);
  }
  free(x); free(R); free(tmp); free(tmp2); free(S); free(Ls); free(RCt); free(K);
  return ll;
}
