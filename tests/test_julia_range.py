"""Julia Base's Float64 range arithmetic behind the systematic-resampling thresholds (reference src/resample.jl:24,
`s = r:(1/M):(bins[N]+r)`; base/twiceprecision.jl is not under the reference tree).

Two independent restatements of the published algorithm — oracle/julia_range.py (Python, exact rationals for the
two_mul error term) and oracle/llpf_oracle.c (C, fma) — must agree bit for bit, and on the fallback path (the one a
53-bit rand() takes) every element must equal the closed form fl(r + fl((i-1)*fl(1/M))) that the CUDA kernels evaluate."""
import math
import random

import numpy as np
import pytest

from oracle import julia_range as J
from oracle import oracle as O


def _cases(n, seed):
    rnd = random.Random(seed)
    for _ in range(n):
        M = rnd.choice([10, 49, 500, 777, 1000, 1234, 4096, 300001, rnd.randint(2, 1 << 22)])
        N = rnd.choice([M, M, rnd.randint(2, 1 << 20)])
        u = rnd.getrandbits(53) * 2.0 ** -53
        total = rnd.choice([1.0, 1.0 - 2.0 ** -53, 1.0 + 2.0 ** -52, 1.0 - 2.0 ** -52, rnd.uniform(0.3, 3.0)])
        yield u * total / N, M, total


def test_rat_known_values():
    assert J.rat(0.1) == (1, 10) and J.rat(0.05) == (1, 20) and J.rat(1 / 3) == (1, 3) and J.rat(0.0) == (0, 1)
    assert J.rat(1 / 49) == (1, 49)          # inv(1/49) != 49 in floating point, the continued fraction still closes
    n, d = J.rat(math.pi)
    assert d != 0 and float(n) / float(d) != math.pi     # an approximation only: the caller's exactness test fails


def test_fallback_path_is_the_closed_form_for_random_draws():
    n_fb = n_rat = 0
    for r, M, total in _cases(100000, 1):
        s = J.julia_thresholds(r, M, total)
        if s.path != "fallback":
            n_rat += 1
            continue
        n_fb += 1
        assert (s.ref_hi, s.ref_lo, s.step_hi, s.step_lo, s.offset) == (r, 0.0, 1.0 / M, 0.0, 1)
        for i in (1, 2, 3, M // 2 + 1, M - 1, M):
            assert s.getindex(i) == J.simple_threshold(r, M, i)
    assert n_fb > 99000          # a 53-bit draw is an exact ratio of integers <= 2^24 only about 1e-3 of the time at M = 10
    assert n_rat < 1000


def test_c_oracle_and_python_restatement_agree_bit_for_bit():
    special = [(0.5 / 10, 10, 1.0), (0.0, 10, 1.0), (0.25 / 777, 777, 1.0), (0.5 / 1234, 1234, 1.0),
               (0.5 / 10, 10, 1.0 - 2.0 ** -53), (0.125 / 37, 100, 0.5), (0.75 / 1000, 1000, 2.0), (0.0, 7, 0.7)]
    cases = special + list(_cases(4000, 2))
    n_rat = 0
    for r, M, total in cases:
        s = J.julia_thresholds(r, M, total)
        R, get = O.julia_range(r, 1.0 / M, total + r)
        assert bool(R.rational) == (s.path == "rational"), (r, M, total)
        assert (R.ref_hi, R.ref_lo, R.step_hi, R.step_lo, R.len, R.offset) == \
               (s.ref_hi, s.ref_lo, s.step_hi, s.step_lo, s.len, s.offset), (r, M, total)
        n_rat += R.rational
        for i in (1, 2, 3, 4, M // 3 + 1, M - 1, M, M + 1):
            assert get(i) == s.getindex(i)
    assert n_rat >= 6


def test_rational_path_differs_from_the_closed_form_only_in_the_last_bit():
    r, M = 0.25 / 777, 777
    s = J.julia_thresholds(r, M, 1.0)
    assert s.path == "rational" and s.len == M + 1
    d = [abs(s.getindex(i) - J.simple_threshold(r, M, i)) / math.ulp(s.getindex(i)) for i in range(1, M + 1)]
    assert 0 < max(d) <= 1.0
    # double-double evaluation == correctly rounded exact rational (start_n + (i-1) step_n)/den
    from fractions import Fraction
    for i in range(1, M + 1):
        exact = Fraction(r) + (i - 1) * Fraction(1, M)
        assert s.getindex(i) == float(exact)


def _walk(we, thr):
    """resample.jl:19-34 in pure Python with thresholds thr(i), i 1-based"""
    N = len(we)
    bins = [we[0]]
    for k in range(1, N):
        bins.append(bins[-1] + we[k])
    j, bo = list(range(1, N + 1)), 0
    for i in range(1, N + 1):
        si = thr(i, bins[-1])
        for b in range(bo, N):
            if si < bins[b]:
                j[i - 1] = b + 1
                bo = b
                break
    return j, bins


def test_rand_zero_uniform_weights_known_answer():
    """resample(uniform we) with rand() == 0.  When the cumulative sum ends at exactly 1.0 Julia takes the rational path
    and s[i] is the correctly rounded (i-1)/M: for N = 5, s[4] = 0.6 < bins[3] = 0.2+0.2+0.2 = 0.6000000000000001
    selects particle 3 for slot 4, where the closed form fl(3*0.2) would select 4.  The C oracle follows Julia."""
    seen_difference = False
    for N in (5, 10, 7, 12, 25, 100, 1000):
        we = np.full(N, 1.0 / N)
        jl, bins = _walk(list(we), lambda i, tot: J.julia_thresholds(0.0, N, tot).getindex(i))
        cf, _ = _walk(list(we), lambda i, tot: J.simple_threshold(0.0, N, i))
        j, b = O.resample_systematic(we, 0.0)
        assert list(j) == jl and list(b) == bins, N
        seen_difference |= jl != cf
    assert seen_difference
    jl5, _ = _walk([0.2] * 5, lambda i, tot: J.julia_thresholds(0.0, 5, tot).getindex(i))
    assert jl5 == [1, 2, 3, 3, 5]      # closed form: [1, 2, 3, 4, 5]
