// llpf_julia_range.h — host side: how Julia Base builds the Float64 range of the systematic-resampling thresholds,
//     s = r:(1/M):(bins[N]+r)                                  (reference src/resample.jl:24)
// base/twiceprecision.jl is not part of the reference tree; this restates its published algorithm (Julia 1.6 - 1.11):
// `(:)(start, step, stop)` first tries to write start, step and stop as exact ratios of integers <= 2^24 (`rat`); if that
// succeeds the range is built in double-double ("TwicePrecision") arithmetic from the integer ratios (`floatrange`),
// otherwise start and step are taken literally.  Either way element i is
//     u = i - offset ; (x_hi, x_lo) = add12(ref.hi, u*step.hi) ; s[i] = x_hi + (x_lo + (u*step.lo + ref.lo))
// which on the literal path (ref.lo = step.lo = 0, offset = 1) is fl(r + fl((i-1)*fl(1/M))) — what the fused engine
// evaluates.  A 53-bit rand() lands on the rational path about 1e-3 of the time at N = 10 and 1e-8 at N = 2^20, and the two
// paths then differ by at most one ulp per threshold; the stand-alone entry llpf_resample_systematic (caller-supplied
// rand(), so 0, 1/2, ... are fair game) follows Julia on both paths: the host takes the decision below, the kernel
// evaluates the general element formula (llpf_engine.cuh `threshold`).  Checked against the oracle's two independent
// restatements (oracle/julia_range.py, oracle/llpf_oracle.c) by the GPU parity tests.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace llpf {

struct RangeTP {          // Julia's StepRangeLen{Float64, TwicePrecision{Float64}, TwicePrecision{Float64}}
  double ref_hi, ref_lo, step_hi, step_lo;
  long long len;
  int offset;             // 1-based index of the reference element
  int rational;           // 1: built by floatrange (double-double), 0: literal start/step
};

namespace jlrange {
inline void rat(double x, long long& num, long long& den) {
  double y = x;
  long long a = 1, d = 1, b = 0, c = 0;
  const double m = 16777216.0;   // maxintfloat(Float32): rat() narrows Float64 -> Float32
  while (std::fabs(y) <= m) {
    const long long f = (long long)y;
    y -= (double)f;
    const long long a2 = f * a + c, b2 = f * b + d;
    c = a; a = a2; d = b; b = b2;
    const long long mx = std::llabs(a) > std::llabs(b) ? std::llabs(a) : std::llabs(b);
    if (!(mx <= (long long)m)) { num = c; den = d; return; }
    if ((double)a / (double)b == x) break;
    y = 1.0 / y;
  }
  num = a; den = b;
}
inline bool isbetween(double a, double x, double b) { return (a <= x && x <= b) || (b <= x && x <= a); }
inline void canonicalize2(double big, double little, double& h, double& l) {
  h = big + little;
  l = (big - h) + little;
}
inline void tp_ratio(long long n, long long d, int nb, double& h, double& l) {
  // TwicePrecision{Float64}(n) / d, then twiceprecision(., nb) when nb >= 0
  const double xhi = (double)n, yhi = (double)d;
  const double hi = xhi / yhi;
  const double uh = hi * yhi;
  const double ul = std::fma(hi, yhi, -uh);
  const double lo = ((((xhi - uh) - ul) + 0.0) - hi * 0.0) / yhi;
  canonicalize2(hi, lo, h, l);
  if (nb >= 0) {
    unsigned long long u;
    std::memcpy(&u, &h, 8);
    u &= (nb >= 64) ? 0ull : (~0ull << nb);
    double h2;
    std::memcpy(&h2, &u, 8);
    l = (h - h2) + l;
    h = h2;
  }
}
inline int nbitslen(long long len, long long offset) {
  if (len < 2) return 0;
  const long long a = offset - 1, b = len - offset;
  const int nb = (int)std::ceil(std::log2((double)(a > b ? a : b))) + 1;
  return nb < 27 ? nb : 27;
}
inline long long gcd(long long a, long long b) {
  while (b) { const long long t = a % b; a = b; b = t; }
  return std::llabs(a);
}
}  // namespace jlrange

inline RangeTP julia_range(double start, double step, double stop) {
  using namespace jlrange;
  RangeTP R;
  long long step_n, step_d, start_n, start_d, stop_n, stop_d;
  rat(step, step_n, step_d);
  if (step_d != 0 && (double)step_n / (double)step_d == step) {
    rat(start, start_n, start_d);
    rat(stop, stop_n, stop_d);
    if (start_d != 0 && stop_d != 0 && (double)start_n / (double)start_d == start &&
        (double)stop_n / (double)stop_d == stop) {
      const long long den = start_d / gcd(start_d, step_d) * step_d;
      const double m = 9007199254740992.0;
      if (den != 0 && std::fabs(start * (double)den) <= m && std::fabs(step * (double)den) <= m &&
          den % start_d == 0 && den % step_d == 0) {
        start_n = (long long)std::nearbyint(start * (double)den);
        step_n = (long long)std::nearbyint(step * (double)den);
        const __int128 q = ((__int128)den * stop_n) / stop_d;
        long long len = (long long)((q - start_n) / step_n) + 1;
        if (len < 0) len = 0;
        if (isbetween(start, start + (double)(len - 1) * step, stop + step / 2) &&
            !isbetween(start, start + (double)len * step, stop)) {
          R.rational = 1; R.len = len;
          long long imin = 1;
          int nb = 0;
          if (!(len < 2 || step_n == 0)) {
            imin = (long long)std::nearbyint(-(double)start_n / (double)step_n + 1);
            if (imin < 1) imin = 1;
            if (imin > len) imin = len;
            nb = nbitslen(len, imin);
          }
          R.offset = (int)imin;
          tp_ratio(start_n + (imin - 1) * step_n, den, -1, R.ref_hi, R.ref_lo);
          tp_ratio(step_n, den, nb, R.step_hi, R.step_lo);
          return R;
        }
      }
    }
  }
  const double lf = (stop - start) / step;
  long long len;
  if (lf < 0) len = 0;
  else if (lf == 0) len = 1;
  else {
    len = (long long)std::nearbyint(lf) + 1;
    const double stop2 = start + (double)(len - 1) * step;
    len -= (start < stop && stop < stop2) + (start > stop && stop > stop2);
  }
  R.rational = 0; R.len = len; R.offset = 1;
  R.ref_hi = start; R.ref_lo = 0.0; R.step_hi = step; R.step_lo = 0.0;
  return R;
}

}  // namespace llpf
