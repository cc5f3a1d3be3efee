"""Host-side bit-level emulation (exact fma via Fractions) of llpf_math.cuh, checked against 50-digit
references: validates the table/coefficients and the reduction logic before spending GPU time."""
import re, struct, sys
from fractions import Fraction as Fr
from decimal import Decimal as D, getcontext
import numpy as np
getcontext().prec = 60
PI = D("3.14159265358979323846264338327950288419716939937510582097494459230781640628620899")
txt = open('lowlevelparticlefilters.jl_b200/csrc/llpf_math_tables.inc').read()
def arr(name):
    m = re.search(name + r'(?:\[[^\]]*\])?\s*=\s*\{?([^;]*?)\}?;', txt, re.S)
    body = re.sub(r'/\*.*?\*/', '', m.group(1))
    return [float(x) for x in body.replace('\n', ' ').split(',') if x.strip()]
def fma(a, b, c): return float(Fr(a) * Fr(b) + Fr(c))
logtab = np.array(arr('c_log_tab')).reshape(-1, 2); ln2hi = arr('c_ln2_hi')[0]; ln2lo = arr('c_ln2_lo')[0]; l1p = arr('c_l1p')
sinc = arr('c_sinpi'); cosc = arr('c_cospi'); extab = arr('c_exp_tab'); einv = arr('c_exp_inv')[0]; ehi = arr('c_exp_hi')[0]; elo = arr('c_exp_lo')[0]; ep = arr('c_exp_p')
def bits(d): return struct.unpack('<Q', struct.pack('<d', d))[0]
def frombits(b): return struct.unpack('<d', struct.pack('<Q', b))[0]
def log_u32(r):
    d = fma(float(r), 2.0, 1.0); b = bits(d); hi = b >> 32; lo = b & 0xffffffff
    e = (hi >> 20) - (1023 + 33); mh = hi & 0xfffff; k = (mh + 0x1000) >> 13; wrap = (k + 64) >> 7; e += wrap
    m = frombits(((mh | (0x3ff00000 - (wrap << 20))) << 32) | lo)
    inv, lc = logtab[k]; rr = fma(m, inv, -1.0)
    p = fma(rr, l1p[5], l1p[4])
    for i in (3, 2, 1, 0): p = fma(rr, p, l1p[i])
    l = fma(rr * rr, p, rr); ef = float(e); hp = fma(ef, ln2hi, lc); return hp + fma(ef, ln2lo, l)
def sincospi_u32(r):
    k = ((r >> 29) + 1) >> 1; ti = (r - (k << 30)) & 0xffffffff; ti = ti - 2**32 if ti >= 2**31 else ti
    t = fma(float(ti), 4.6566128730773926e-10, 2.3283064365386963e-10); t2 = t * t
    sp = fma(t2, sinc[8], sinc[7]); cp = fma(t2, cosc[8], cosc[7])
    for i in range(6, -1, -1): sp = fma(t2, sp, sinc[i]); cp = fma(t2, cp, cosc[i])
    st = t * sp; swap = k & 1; ss = cp if swap else st; cc = st if swap else cp
    if k & 2: ss = -ss
    if (k + 1) & 2: cc = -cc
    return ss, cc
def exp_nonpos(x):
    magic = 6755399441055744.0; xs = max(x, -720.0); nm = fma(xs, einv, magic)
    n = bits(nm) & 0xffffffff; n = n - 2**32 if n >= 2**31 else n; nf = nm - magic
    r = fma(nf, -ehi, xs); r = fma(nf, -elo, r)
    p = fma(r, ep[4], ep[3])
    for i in (2, 1, 0): p = fma(r, p, ep[i])
    q = fma(r * r, p, r); tj = extab[n & 63]; v = fma(tj, q, tj); sh = n >> 6
    b = bits(v); hi = ((b >> 32) + (sh << 20)) & 0xffffffff
    res = frombits((hi << 32) | (b & 0xffffffff))
    return 0.0 if x < -708.0 else res
def dsin(x):  # Decimal sin/cos by Taylor
    x = D(x); s = D(0); term = x; k = 0
    while abs(term) > D(10) ** -55:
        s += term; k += 1; term = -term * x * x / ((2 * k) * (2 * k + 1))
    return s
def dcos(x):
    x = D(x); s = D(0); term = D(1); k = 0
    while abs(term) > D(10) ** -55:
        s += term; k += 1; term = -term * x * x / ((2 * k - 1) * (2 * k))
    return s
rng = np.random.default_rng(0)
rs = [int(v) for v in rng.integers(0, 2**32, size=2000)] + [0, 1, 2, 2**32 - 1, 2**32 - 2, 2**31, 2**31 - 1, 2**32 - 2**10, 2**29, 2**29 - 1, 3 * 2**29, 7 * 2**29 - 1, 7 * 2**29]
mx = 0
for r in rs:
    ex = ((D(r) + D('0.5')) / D(2**32)).ln(); got = log_u32(r); mx = max(mx, float(abs((D(got) - ex) / ex)))
print('log  max rel err %.3e (%.2f ulp)' % (mx, mx / 1.11e-16))
ms = mc = 0
for r in rs:
    a = (D(r) + D('0.5')) / D(2**31); red = a if a <= 1 else a - 2  # keep the Taylor reference fast
    es, ec = dsin(PI * red), dcos(PI * red); s, c = sincospi_u32(r)
    ms = max(ms, float(abs(D(s) - es))); mc = max(mc, float(abs(D(c) - ec)))
print('sincospi max abs err %.3e %.3e' % (ms, mc))
me = 0
for x in list(-rng.random(1500) * 50) + list(-rng.random(300) * 700) + [0.0, -1e-300, -1e-17, -0.0054, -707.9, -708.5, -1000.0, float('-inf')]:
    got = exp_nonpos(float(x))
    if x < -708: assert got == 0.0; continue
    ex = D(float(x)).exp(); me = max(me, float(abs((D(got) - ex) / ex)))
print('exp  max rel err %.3e (%.2f ulp)' % (me, me / 1.11e-16))
