"""exact_scan.py — TEST INFRASTRUCTURE / design groundwork: a parallel-friendly scan that reproduces the reference's serial
Float64 `cumsum` (resample.jl:18-21: bins[i] = bins[i-1] + we[i], left to right) BIT FOR BIT.

Why: the product's FAST scan is an exact fixed-point integer scan (2^-62 units).  It is order-independent and bit-identical
on 1 or 8 GPUs, but its bins differ from the serial f64 cumsum by rounding, so at N = 2^20 a threshold occasionally falls on
the other side of a bin edge and the realisation diverges from the oracle's.  This file shows that the serial
result itself can be computed with O(N) parallel work plus a sequential pass over a few dozen "special" elements, and checks
it on adversarial inputs (tests/test_exact_scan.py).  Not used by the product yet.

Idea.  While the running sum B stays inside one binade [2^e, 2^(e+1)), every B is a multiple of u = 2^(e-52), and
    fl(B + w) = B + RN(w / u) * u                      (round to nearest; B + w < 2^(e+1))
unless w/u lies exactly half-way between two integers (then the result depends on the parity of B/u).  So for a REGULAR element —
the binade of the running sum before and after it is known for certain, and w/u is not a tie — the increment d_i = RN(w_i/u)*u
does not depend on B at all.  Everything else is SPECIAL: the elements at which the sum may change binade, exact ties, and
elements whose binade cannot be certified from the approximate prefix.  Then
    B_i = B_s + (sum of d_k over the regular k in (s, i]),   s = last special element <= i,
where the partial sums are exact (all terms are multiples of one u and the results are representable), and the specials are
resolved left to right with ONE f64 addition each: B_s = fl(B_{s-1} + w_s).
Certification uses only what a parallel pass has: the exact fixed-point prefix A_i of the quantised weights (what the product's
scan already computes) and the bound |B_i - A_i| <= (i+1) * 2^-53 * A_i + (i+1) * 2^-63.
"""
import math

import numpy as np

FIX_BITS = 62


def serial_cumsum(w):
    """the reference: left-to-right Float64 adds"""
    out = np.empty(len(w))
    b = 0.0
    for i, v in enumerate(w):
        b = b + float(v)
        out[i] = b
    return out


def _binade(x):
    """e with 2^e <= x < 2^(e+1) (x > 0)"""
    return math.frexp(x)[1] - 1


def exact_scan(w, return_stats=False):
    """Serial-equivalent inclusive scan of the non-negative doubles w.  Phases 1-3 and 5 are element-parallel (written as numpy
    vector operations / integer scans); phase 4 walks the special elements only."""
    w = np.asarray(w, dtype=np.float64)
    n = w.size
    if n == 0:
        return (w.copy(), dict(special=0)) if return_stats else w.copy()
    # phase 1: exact fixed-point prefix of the quantised weights (Python ints here; u64 on the device)
    q = [int(round(float(v) * 2.0 ** FIX_BITS)) for v in w]          # |q_i / 2^62 - w_i| <= 2^-63
    A = np.empty(n)
    acc = 0
    Aint = [0] * n
    for i in range(n):
        acc += q[i]
        Aint[i] = acc
        A[i] = acc / 2.0 ** FIX_BITS
    # phase 2: certify the binade of the running sum around every element
    idx = np.arange(1, n + 1, dtype=np.float64)
    delta = idx * (2.0 ** -53) * A + idx * 2.0 ** -63               # bound on |B_i - A_i|
    lo = np.concatenate(([0.0], A[:-1] - delta[:-1]))               # lower bound of B_{i-1}
    hi = A + delta                                                   # upper bound of B_i
    special = np.zeros(n, dtype=bool)
    e_of = np.zeros(n, dtype=np.int64)
    d_units = [0] * n                                                # RN(w_i / u) for regular elements
    for i in range(n):
        if not (lo[i] > 0.0):
            special[i] = True
            continue
        e = _binade(lo[i])
        if not (hi[i] < 2.0 ** (e + 1)) or lo[i] <= 2.0 ** e:        # may leave the binade (or sits on its lower edge)
            special[i] = True
            continue
        # phase 3: the increment in units of u = 2^(e-52): exact integer arithmetic on the double's mantissa
        m, ex = math.frexp(float(w[i]))                              # w = m * 2^ex, 0.5 <= m < 1
        mi = int(m * 2.0 ** 53)                                      # 53-bit integer mantissa, w = mi * 2^(ex-53)
        sh = (e - 52) - (ex - 53)                                    # w / u = mi / 2^sh
        if sh <= 0:
            r = mi << (-sh)
        else:
            r, rem = mi >> sh, mi & ((1 << sh) - 1)
            half = 1 << (sh - 1)
            if rem == half:                                          # exact tie: depends on the parity of B/u
                special[i] = True
                continue
            if rem > half:
                r += 1
        e_of[i] = e
        d_units[i] = r
    # phase 4: resolve the specials left to right; between two specials the sum advances by an exact integer number of units
    B = np.empty(n)
    spec_idx = np.flatnonzero(special)
    prev_special, b_prev = -1, 0.0
    seg_base = {}                                                    # first regular index after a special -> B at that special
    for s in list(spec_idx) + [n]:
        # regular elements (prev_special, s): one segment, one binade
        if s - prev_special > 1:
            first = prev_special + 1
            u = 2.0 ** (int(e_of[first]) - 52)
            tot = 0
            for k in range(first, s):                                # (an integer scan on the device)
                tot += d_units[k]
                B[k] = b_prev + tot * u                              # exact: representable by construction
            b_prev = B[s - 1]
        if s < n:
            b_prev = b_prev + float(w[s])                            # ONE f64 addition: the reference's own operation
            B[s] = b_prev
            prev_special = s
    if return_stats:
        return B, dict(special=int(special.sum()), n=n)
    return B


# ---------------------------------------------------------------------------------------------------------------------
# The same algorithm in the decomposition a GPU would run (blocks of `chunk` consecutive elements, nothing but block-local
# work, scans over per-block aggregates, and ONE sequential walk over the special elements):
#   A  fixed-point block scan -> block totals -> exclusive block offsets            (what resample_indices does today)
#   B  per element: certify the binade from the global fixed-point prefix, r_i = RN(w_i / u) or "special"
#   C  segmented integer scan of r (restarting after every special): block-local scan + a segmented scan over the block
#      aggregates (carry = units accumulated since the last special, which is in the same binade by construction);
#      alongside, a max-scan of "index of the last special <= i"
#   D  one thread walks the specials in index order: B_before = B_(previous special) + u * (units accumulated before s),
#      B_s = fl(B_before + w_s)                                                      (one f64 add per special)
#   E  per element: B_i = B_(last special <= i) + u_i * (units since it)
# ---------------------------------------------------------------------------------------------------------------------
def exact_scan_blocked(w, chunk=3544, return_stats=False):
    w = np.asarray(w, dtype=np.float64)
    n = w.size
    if n == 0:
        return (w.copy(), dict(special=0)) if return_stats else w.copy()
    nb = (n + chunk - 1) // chunk
    blocks = [(b * chunk, min(n, (b + 1) * chunk)) for b in range(nb)]
    # ---- A
    q = [int(round(float(v) * 2.0 ** FIX_BITS)) for v in w]
    loc = [0] * n                         # block-local inclusive prefix
    tot = [0] * nb
    for b, (s, e) in enumerate(blocks):
        acc = 0
        for i in range(s, e):
            acc += q[i]
            loc[i] = acc
        tot[b] = acc
    off = [0] * nb
    for b in range(1, nb):
        off[b] = off[b - 1] + tot[b - 1]
    # ---- B (element-parallel; needs only off[block], loc[i], loc[i-1], w[i], i)
    special = [False] * n
    e_of = [0] * n
    r = [0] * n
    for b, (s, e) in enumerate(blocks):
        for i in range(s, e):
            a_hi = (off[b] + loc[i]) / 2.0 ** FIX_BITS
            a_lo = (off[b] + loc[i] - q[i]) / 2.0 ** FIX_BITS
            d_hi = (i + 1) * 2.0 ** -53 * a_hi + (i + 1) * 2.0 ** -63
            d_lo = i * 2.0 ** -53 * a_lo + i * 2.0 ** -63
            lo, hi = a_lo - d_lo, a_hi + d_hi
            if not (lo > 0.0):
                special[i] = True
                continue
            eb = _binade(lo)
            if not (hi < 2.0 ** (eb + 1)) or lo <= 2.0 ** eb:
                special[i] = True
                continue
            m, ex = math.frexp(float(w[i]))
            mi = int(m * 2.0 ** 53)
            sh = (eb - 52) - (ex - 53)
            if sh <= 0:
                ri = mi << (-sh)
            else:
                ri, rem = mi >> sh, mi & ((1 << sh) - 1)
                half = 1 << (sh - 1)
                if rem == half:
                    special[i] = True
                    continue
                if rem > half:
                    ri += 1
            e_of[i], r[i] = eb, ri
    # ---- C: segmented scan, block-local part
    seg = [0] * n                         # units since the last special at or before i (0 at a special), block-local
    last = [-1] * n                       # index of the last special <= i inside the block, -1 if none
    agg_units = [0] * nb                  # units after the block's last special (or of the whole block if it has none)
    agg_has = [False] * nb
    agg_last = [-1] * nb
    for b, (s, e) in enumerate(blocks):
        acc, ls = 0, -1
        for i in range(s, e):
            if special[i]:
                acc, ls = 0, i
            else:
                acc += r[i]
            seg[i], last[i] = acc, ls
        agg_units[b], agg_has[b], agg_last[b] = acc, ls >= 0, ls
    # segmented scan over the block aggregates: what every block inherits from its predecessors
    carry_units = [0] * nb
    carry_last = [-1] * nb
    for b in range(1, nb):
        if agg_has[b - 1]:
            carry_units[b], carry_last[b] = agg_units[b - 1], agg_last[b - 1]
        else:
            carry_units[b], carry_last[b] = carry_units[b - 1] + agg_units[b - 1], carry_last[b - 1]
    units = [0] * n                       # units since the globally last special <= i
    glast = [-1] * n
    for b, (s, e) in enumerate(blocks):
        for i in range(s, e):
            if last[i] >= 0:
                units[i], glast[i] = seg[i], last[i]
            else:
                units[i], glast[i] = carry_units[b] + seg[i], carry_last[b]
    # ---- D: the sequential walk
    spec_idx = [i for i in range(n) if special[i]]
    Bs = {}
    b_prev = 0.0                          # B before the first element
    for s_ in spec_idx:
        if s_ > 0 and not special[s_ - 1]:
            base = Bs[glast[s_ - 1]] if glast[s_ - 1] >= 0 else 0.0
            b_before = base + units[s_ - 1] * 2.0 ** (e_of[s_ - 1] - 52)
        else:
            b_before = Bs[s_ - 1] if s_ > 0 else 0.0
        Bs[s_] = b_before + float(w[s_])
    # ---- E
    out = np.empty(n)
    for i in range(n):
        if special[i]:
            out[i] = Bs[i]
        else:
            base = Bs[glast[i]] if glast[i] >= 0 else 0.0
            out[i] = base + units[i] * 2.0 ** (e_of[i] - 52)
    if return_stats:
        return out, dict(special=len(spec_idx), n=n, blocks=nb)
    return out
