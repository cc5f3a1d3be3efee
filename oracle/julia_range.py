"""julia_range.py — TEST INFRASTRUCTURE (oracle side), NOT PRODUCT CODE.

A literal Python restatement of how Julia Base builds and indexes the Float64 range

    s = r:(1/M):(bins[N]+r)                                   (reference: src/resample.jl:24)

whose elements `s[i]` are the systematic-resampling thresholds compared against `bins` (resample.jl:28).
The arithmetic lives in Julia Base (`base/twiceprecision.jl`), which is NOT under /root/reference and cannot be run here
(no julia in the image): this file restates the published algorithm of Julia 1.6 - 1.11 (the function bodies below have
been stable across those releases), function by function, so that the oracle's threshold formula is a *derivation*
from that algorithm and not a recollection of its result:

    rat(x)                       continued-fraction rational approximation, denominators bounded by maxintfloat(Float32)
    (:)(start, step, stop)       "nice rational" path -> floatrange(...) ; otherwise the literal fallback
    floatrange / steprangelen_hp / TwicePrecision{Float64}((n, d)[, nb]) / twiceprecision / truncbits / nbitslen
    add12 / mul12 / canonicalize2 / TwicePrecision division
    unsafe_getindex(r::StepRangeLen{T,<:TwicePrecision,<:TwicePrecision}, i)

Python floats are IEEE-754 binary64 and every `+ - * /` below is one correctly rounded operation (CPython never
contracts to FMA), i.e. the same arithmetic Julia performs.  `mul12` needs the exact product error: taken with
exact rational arithmetic (fractions.Fraction), which equals Julia's fma-based two_mul.

Result (tests/test_julia_range.py):
  * fallback path (start or stop not an exact small rational — the case for r = rand()*bins[end]/N with a 53-bit
    rand(): both `r` and `bins[N]+r` would have to be ratios of integers <= 2^24):
        s[i] == fl(r + fl((i-1)*fl(1/M)))        for every i       (the formula of oracle/llpf_oracle.c and the kernel)
  * rational path (e.g. rand() == 0, or dyadic test inputs such as u = 0.5 with weights summing to exactly 1):
        s[i] is the double-double evaluation (start_n + (i-1)*step_n)/den, which can differ from the formula above
        in the last bit when M is not a power of two.  `julia_thresholds` returns whichever Julia would produce;
        the C oracle and the stand-alone CUDA entry take the same decision (orc_julia_range / llpf_api.cu).
"""
import math
import struct
from fractions import Fraction

MAXINTFLOAT_F32 = 16777216          # maxintfloat(Float32, Int): narrow(Float64) == Float32 in rat()
MAXINTFLOAT_F64 = 9007199254740992  # maxintfloat(Float64, Int)


def _trunc_int(y):
    return int(y)                   # trunc(Int, y)


def rat(x):
    """base/twiceprecision.jl `rat(x)`: (numerator, denominator) of a continued-fraction approximation."""
    y = x
    a = d = 1
    b = c = 0
    m = MAXINTFLOAT_F32
    while abs(y) <= m:
        f = _trunc_int(y)
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if not (max(abs(a), abs(b)) <= m):
            return c, d
        if b != 0 and float(a) / float(b) == x:
            break
        if y == 0.0:
            # inv(0.0) == Inf: the loop condition `abs(y) <= m` fails on the next test
            y = math.inf
        else:
            y = 1.0 / y
    return a, b


def _rat_exact(n, d, x):
    return d != 0 and float(n) / float(d) == x


def _isbetween(a, x, b):
    return a <= x <= b or b <= x <= a


def _round_int(v):
    """round(Int, v): RoundNearest, ties to even (Julia's default for round(Int, ::Float64))."""
    return int(round(v))            # Python's round() on floats is also ties-to-even


# ---- TwicePrecision helpers ---------------------------------------------------------------------------------------
def canonicalize2(big, little):
    h = big + little
    return h, (big - h) + little


def add12(x, y):
    if abs(y) > abs(x):
        x, y = y, x
    return canonicalize2(x, y)


def mul12(x, y):
    h = x * y
    if not math.isfinite(h):
        return h, h
    err = Fraction(x) * Fraction(y) - Fraction(h)     # exact: representable in binary64 (two_mul / fma)
    return h, float(err)


def _bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def _from_bits(u):
    return struct.unpack("<d", struct.pack("<Q", u & 0xFFFFFFFFFFFFFFFF))[0]


def truncbits(x, nb):
    return _from_bits(_bits(x) & ((0xFFFFFFFFFFFFFFFF << nb) & 0xFFFFFFFFFFFFFFFF))


def nbitslen(length, offset):
    """nbitslen(Float64, len, offset) = min(cld(53, 2), nbitslen(len, offset))"""
    if length < 2:
        nb = 0
    else:
        nb = math.ceil(math.log2(max(offset - 1, length - offset))) + 1
    return min(27, nb)


def tp_div(xhi, xlo, yhi, ylo):
    """/(x::TwicePrecision, y::TwicePrecision)"""
    hi = xhi / yhi
    uh, ul = mul12(hi, yhi)
    lo = ((((xhi - uh) - ul) + xlo) - hi * ylo) / yhi
    return canonicalize2(hi, lo)


def tp_from_ratio(n, d, nb=None):
    """TwicePrecision{Float64}((n, d)) = TwicePrecision{Float64}(n) / d ; with nb: twiceprecision(..., nb)."""
    nhi = float(n)
    nlo = float(n - int(nhi))
    dhi = float(d)
    hi, lo = tp_div(nhi, nlo, dhi, float(d - int(dhi)))
    if nb is not None:
        h2 = truncbits(hi, nb)
        hi, lo = h2, (hi - h2) + lo
    return hi, lo


class StepRangeLenTP:
    """StepRangeLen{Float64, TwicePrecision{Float64}, TwicePrecision{Float64}}"""

    def __init__(self, ref, step, length, offset, path):
        self.ref_hi, self.ref_lo = ref
        self.step_hi, self.step_lo = step
        self.len = int(length)
        self.offset = int(offset)
        self.path = path            # "rational" | "fallback"

    def getindex(self, i):
        """unsafe_getindex(r, i), i 1-based (no bounds check: resample.jl:27 indexes under @inbounds)."""
        u = i - self.offset
        shift_hi, shift_lo = u * self.step_hi, u * self.step_lo
        x_hi, x_lo = add12(self.ref_hi, shift_hi)
        return x_hi + (x_lo + (shift_lo + self.ref_lo))


def floatrange(start_n, step_n, length, den):
    if length < 2 or step_n == 0:
        return StepRangeLenTP(tp_from_ratio(start_n, den), tp_from_ratio(step_n, den, 0), length, 1, "rational")
    # index of the smallest-magnitude value
    imin = min(max(_round_int(-start_n / step_n + 1), 1), length)
    ref_n = start_n + (imin - 1) * step_n
    nb = nbitslen(length, imin)
    return StepRangeLenTP(tp_from_ratio(ref_n, den), tp_from_ratio(step_n, den, nb), length, imin, "rational")


def colon(start, step, stop):
    """(:)(start::Float64, step::Float64, stop::Float64)"""
    if step == 0:
        raise ValueError("range step cannot be zero")
    step_n, step_d = rat(step)
    if _rat_exact(step_n, step_d, step):
        start_n, start_d = rat(start)
        stop_n, stop_d = rat(stop)
        if _rat_exact(start_n, start_d, start) and _rat_exact(stop_n, stop_d, stop):
            den = start_d * step_d // math.gcd(start_d, step_d)          # lcm_unchecked
            m = MAXINTFLOAT_F64
            if den != 0 and abs(start * den) <= m and abs(step * den) <= m and den % start_d == 0 and den % step_d == 0:
                start_n = _round_int(start * den)
                step_n = _round_int(step * den)
                # len = max(0, Int(div(den*stop_n, stop_d) - start_n) ÷ step_n + 1)
                length = max(0, _tdiv(_tdiv(den * stop_n, stop_d) - start_n, step_n) + 1)
                if _isbetween(start, start + (length - 1) * step, stop + step / 2) and \
                        not _isbetween(start, start + length * step, stop):
                    return floatrange(start_n, step_n, length, den)
    lf = (stop - start) / step
    if lf < 0:
        length = 0
    elif lf == 0:
        length = 1
    else:
        length = _round_int(lf) + 1
        stop2 = start + (length - 1) * step
        length -= int(start < stop < stop2) + int(start > stop > stop2)
    # steprangelen_hp(Float64, start, step, 0, len, 1): ref = (start, 0), step = (step, 0) (nb = 0 truncates nothing)
    return StepRangeLenTP((start, 0.0), (step, 0.0), length, 1, "fallback")


def _tdiv(a, b):
    """Julia `div` / `÷` on Int: truncation toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def julia_thresholds(r, M, total):
    """The StepRangeLen `r:(1/M):(total+r)` of resample.jl:24."""
    return colon(r, 1.0 / M, total + r)


def simple_threshold(r, M, i):
    """fl(r + fl((i-1)*fl(1/M))), i 1-based — the fallback-path closed form used by the C oracle and the CUDA kernels."""
    step = 1.0 / M
    prod = (i - 1) * step
    return r + prod
