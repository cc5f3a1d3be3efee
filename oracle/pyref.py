"""pyref.py — TEST INFRASTRUCTURE: a second, independent restatement of the reference's particle-filter path.

Written in pure Python straight from the Julia sources (file:line cited per function, paths relative to the reference
root), NOT from oracle/llpf_oracle.c: the two restatements are compared bit for bit in tests/test_pyref_twin.py
(SURVEY §7.1 asks for this twin).  Where they agree, a reading error would have had to be made twice, independently.

Only things that are *ours* are shared by construction: the counter-based RNG contract of DESIGN.md §5 (Philox4x32-10,
Box-Muller) — re-implemented here from its specification — and the conventions for third-party arithmetic the reference
does not pin (PDMats `invquad` through the Cholesky factor, StaticArrays left-to-right mat-vec, Base's pairwise `sum`).

Python floats are IEEE binary64, each operator is one rounded operation, `math.exp/log/log1p/sin/cos/sqrt` are the
platform libm — the same functions the C oracle calls.  Pure-Python loops: small N only.
"""
import math

from . import julia_range as JR

ST_INIT, ST_DYN, ST_RESAMPLE, ST_STRAT, ST_RESID = 0, 1, 2, 3, 4
M32 = 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------
# RNG contract — ours, not the reference's
# ---------------------------------------------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def rng_block(seed, epoch, stream, t, i, blk=0):
    key = (seed & M32, (seed >> 32) & M32)
    ctr = (i & M32, (blk + (((i >> 32) & M32) << 16)) & M32, t & M32, (stream | ((epoch & M32) << 8)) & M32)
    return philox4x32_10(ctr, key)


def uniform53(seed, epoch, stream, t, i):
    r = rng_block(seed, epoch, stream, t, i)
    return float(((r[0] << 32) | r[1]) >> 11) * 2.0 ** -53


def _normal_pair(ra, rb):
    u1 = (ra + 0.5) * 2.0 ** -32
    u2 = (rb + 0.5) * 2.0 ** -32
    rad = math.sqrt(-2.0 * math.log(u1))
    a = 2.0 * u2                               # angle / pi in (0, 2)
    q = round(2.0 * a)                         # nearest quadrant boundary (ties to even, like nearbyint)
    r = a - 0.5 * q
    sr, cr = math.sin(math.pi * r), math.cos(math.pi * r)
    s, c = ((sr, cr), (cr, -sr), (-sr, -cr), (-cr, sr))[int(q) & 3]
    return rad * c, rad * s


def normals(seed, epoch, stream, t, i, n):
    z = []
    for b in range((n + 3) // 4):
        r = rng_block(seed, epoch, stream, t, i, b)
        z.extend(_normal_pair(r[0], r[1]))
        z.extend(_normal_pair(r[2], r[3]))
    return z[:n]


# ---------------------------------------------------------------------------------------------------------------
# small dense algebra (matrices are lists of rows)
# ---------------------------------------------------------------------------------------------------------------
def cholesky_lower(S):
    n = len(S)
    L = [[0.0] * n for _ in range(n)]
    for j in range(n):
        d = S[j][j]
        for k in range(j):
            d -= L[j][k] * L[j][k]
        if not d > 0.0:
            raise ValueError("not positive definite")
        d = math.sqrt(d)
        L[j][j] = d
        for i in range(j + 1, n):
            v = S[i][j]
            for k in range(j):
                v -= L[i][k] * L[j][k]
            L[i][j] = v / d
    return L


def matvec(Mx, v):
    """StaticArrays `M*v`: out[r] = M[r,1]*v[1] + M[r,2]*v[2] + ... left to right"""
    out = []
    for row in Mx:
        acc = row[0] * v[0]
        for c in range(1, len(v)):
            acc = acc + row[c] * v[c]
        out.append(acc)
    return out


def lower_times(L, z):
    """cholesky(Sigma).L * z   (utils.jl:260-268)"""
    return [_dot_lower(L[r], z, r) for r in range(len(z))]


def _dot_lower(row, z, r):
    acc = 0.0
    for c in range(r + 1):
        acc += row[c] * z[c]
    return acc


def pairwise_sum(a, lo=0, hi=None):
    """Base.sum(::Vector{Float64}) = mapreduce_impl with pairwise blocks of 1024 (base/reduce.jl)"""
    if hi is None:
        hi = len(a)
    n = hi - lo
    if n <= 0:
        return 0.0
    if n <= 1024:
        s = a[lo]
        for k in range(lo + 1, hi):
            s += a[k]
        return s
    mid = lo + ((n - 1) >> 1) + 1
    return pairwise_sum(a, lo, mid) + pairwise_sum(a, mid, hi)


# ---------------------------------------------------------------------------------------------------------------
# weight numerics — src/utils.jl
# ---------------------------------------------------------------------------------------------------------------
def findmax(w):
    m, k = w[0], 0
    for i in range(1, len(w)):
        if w[i] > m:
            m, k = w[i], i
    return m, k


def sum_all_but(w, i):                         # utils.jl:66-71
    w[i] -= 1
    s = pairwise_sum(w)
    w[i] += 1
    return s


def logsumexp(w, we):                          # utils.jl:18-27 (in place), returns ll
    offset, maxind = findmax(w)
    for i in range(len(w)):
        w[i] -= offset
    for i in range(len(w)):
        we[i] = math.exp(w[i])                 # exp_map! utils.jl:3-7 (SLEEFPirates.exp ~ libm to 1 ulp)
    s = sum_all_but(we, maxind)
    inv = 1 / (s + 1)
    for i in range(len(w)):
        we[i] *= inv
    l1p = math.log1p(s)
    for i in range(len(w)):
        w[i] -= l1p
    return l1p + offset


def expnormalize1(w):                          # utils.jl:57-63
    offset, maxind = findmax(w)
    for i in range(len(w)):
        w[i] -= offset
    for i in range(len(w)):
        w[i] = math.exp(w[i])
    s = sum_all_but(w, maxind)
    inv = 1 / (s + 1)
    for i in range(len(w)):
        w[i] *= inv


def effective_particles(we):                   # resample.jl:1-2   1/sum(abs2, we)
    return 1 / pairwise_sum([v * v for v in we])


# ---------------------------------------------------------------------------------------------------------------
# resampling — src/resample.jl
# ---------------------------------------------------------------------------------------------------------------
def _cumsum(we, bins):
    bins[0] = we[0]
    for i in range(1, len(we)):
        bins[i] = bins[i - 1] + we[i]


def resample_systematic(we, j, bins, u01, M=None):        # resample.jl:17-36 ; rand() of :23 supplied
    N = len(we)
    M = N if M is None else M
    _cumsum(we, bins)
    r = u01 * bins[-1] / N
    s = JR.colon(r, 1 / M, bins[N - 1] + r)
    bo = 0
    for i in range(M):
        si = s.getindex(i + 1)
        for b in range(bo, N):
            if si < bins[b]:
                j[i] = b + 1
                bo = b
                break
    return j


def resample_stratified(we, j, bins, us, M=None):         # resample.jl:38-61 ; rand() of :49 supplied per slot
    N = len(we)
    M = N if M is None else M
    _cumsum(we, bins)
    bo = 0
    for i in range(M):
        u = (i + us[i]) / M * bins[N - 1]
        for b in range(bo, N):
            if u < bins[b]:
                j[i] = b + 1
                bo = b
                break
    return j


def resample_residual(we, j, bins, us, M=None):           # resample.jl:63-117 ; rand() of :106 supplied in draw order
    N = len(we)
    M = N if M is None else M
    wsum = 0.0
    for v in we:
        wsum += v
    inv = 1 / wsum
    num = 0
    for i in range(N):
        nw = we[i] * inv * M
        cnt = math.floor(nw)
        bins[i] = nw - cnt
        for _ in range(cnt):
            j[num] = i + 1
            num += 1
    if num == M:
        return j
    rsum = 0.0
    for v in bins:
        rsum += v
    inv_r = 1 / rsum
    for i in range(N):
        bins[i] *= inv_r
    for i in range(1, N):
        bins[i] += bins[i - 1]
    k = 0
    for m in range(num, M):
        u = us[k]
        k += 1
        for i in range(N):
            if u < bins[i]:
                j[m] = i + 1
                break
    return j


# ---------------------------------------------------------------------------------------------------------------
# models (descriptors of include/llpf.h): linear-Gaussian and the quadtank of examples/example_quadtank.jl
# ---------------------------------------------------------------------------------------------------------------
class Model:
    def __init__(self, C, R1, R2, mu0, Sigma0, A=None, B=None, quadtank=None):
        tl = lambda M_: [list(map(float, row)) for row in M_]   # noqa: E731
        self.A = tl(A) if A is not None else None
        self.B = tl(B) if B is not None and len(B[0]) > 0 else None
        self.C = tl(C)
        self.nx, self.ny = len(self.C[0]), len(self.C)
        self.L1 = cholesky_lower(tl(R1))
        self.L2 = cholesky_lower(tl(R2))
        self.L0 = cholesky_lower(tl(Sigma0))
        self.mu0 = list(map(float, mu0))
        ld = 0.0
        for i in range(self.ny):
            ld += math.log(self.L2[i][i])
        ld *= 2
        self.c0 = -(self.ny * math.log(2 * math.pi) + ld) / 2          # mvnormal_c0 utils.jl:254-257
        self.qt = quadtank          # dict(p, t_switch, a1_factor, Ts, supersample) or None

    # dynamics(x,u,p,t) without noise
    def dynamics(self, x, u, t):
        if self.qt is not None:
            return self._rk4(x, u, t)
        ax = matvec(self.A, x)                                         # A*x .+ B*u  example_lineargaussian.jl:27
        if self.B is not None:
            bu = matvec(self.B, u)
            return [ax[r] + bu[r] for r in range(self.nx)]
        return ax

    def _quadtank(self, h, u, t):                                      # example_quadtank.jl:91-106 (+ :15-17)
        kc, k1, k2, A_, a, gam = self.qt["p"]
        g = 9.81
        a1 = a
        if t > self.qt["t_switch"]:
            a1 *= self.qt["a1_factor"]
        ssqrt = lambda v: math.sqrt(max(v, 0.0) + 1e-3)                # noqa: E731
        tg = 2 * g
        return [
            -a1 / A_ * ssqrt(tg * h[0]) + a / A_ * ssqrt(tg * h[2]) + gam * k1 / A_ * u[0],
            -a / A_ * ssqrt(tg * h[1]) + a / A_ * ssqrt(tg * h[3]) + gam * k2 / A_ * u[1],
            -a / A_ * ssqrt(tg * h[2]) + (1 - gam) * k2 / A_ * u[1],
            -a / A_ * ssqrt(tg * h[3]) + (1 - gam) * k1 / A_ * u[0],
        ]

    def _rk4(self, x0, u, t):                                          # utils.jl:220-237
        ss = self.qt["supersample"]
        Ts = self.qt["Ts"] / ss
        x = list(x0)
        n = len(x)
        for _ in range(ss):
            f1 = self._quadtank(x, u, t)
            f2 = self._quadtank([x[i] + Ts / 2 * f1[i] for i in range(n)], u, t + Ts / 2)
            f3 = self._quadtank([x[i] + Ts / 2 * f2[i] for i in range(n)], u, t + Ts / 2)
            f4 = self._quadtank([x[i] + Ts * f3[i] for i in range(n)], u, t + Ts)
            x = [x[i] + Ts / 6 * (f1[i] + 2 * f2[i] + 2 * f3[i] + f4[i]) for i in range(n)]
            t += Ts
        return x

    def logpdf_meas(self, r):                                          # utils.jl:252 ; invquad = |L \\ r|^2 (PDMats)
        v = []
        q = 0.0
        for i in range(self.ny):
            acc = r[i]
            for k in range(i):
                acc -= self.L2[i][k] * v[k]
            v.append(acc / self.L2[i][i])
            q += v[i] * v[i]
        return self.c0 - q / 2


# ---------------------------------------------------------------------------------------------------------------
# the filters — src/PFtypes.jl, src/filtering.jl, src/smoothing.jl
# ---------------------------------------------------------------------------------------------------------------
PF, ADVANCED, AUX, AUX_ADVANCED = 0, 1, 2, 3


class Filter:
    def __init__(self, model, N, kind=PF, resampling=0, resample_threshold=0.1, Ts=1.0, seed=0, inject=None, record=None):
        """inject: dict(x0=[N][nx], noise=[K][N][nx], u_res=[K][...]) — variates recorded from the reference itself
        (julia/dump_golden.jl) consumed in place of the counter-based streams; K counts predict! calls since reset!.
        record: a dict that receives the same three tables (what this run consumed)."""
        self.inject, self.record = inject, record
        self.k = 0
        self.m, self.N, self.kind = model, N, kind
        self.resampling, self.thr, self.Ts, self.seed = resampling, resample_threshold, Ts, seed
        self.epoch = 0
        nx = model.nx
        self.x = [[0.0] * nx for _ in range(N)]
        self.xprev = [[0.0] * nx for _ in range(N)]
        self.w = [math.log(1 / N)] * N                                 # PFtypes.jl:68
        self.we = [1 / N] * N
        self.j = list(range(1, N + 1))                                 # collect(1:N)  PFtypes.jl:70
        self.bins = [0.0] * N
        self.t = 0
        self.nres = 0

    @property
    def aux(self):
        return self.kind in (AUX, AUX_ADVANCED)

    def reset(self, epoch=0):                                          # filtering.jl:4-14
        self.epoch = epoch
        self.k = 0
        m, N = self.m, self.N
        for i in range(N):
            if self.inject is not None:
                self.xprev[i] = list(map(float, self.inject["x0"][i]))
            else:
                z = normals(self.seed, epoch, ST_INIT, 0, i, m.nx)
                lz = lower_times(m.L0, z)
                self.xprev[i] = [m.mu0[r] + lz[r] for r in range(m.nx)]    # rand(rng, d0) = mu + L z  utils.jl:260
            self.x[i] = list(self.xprev[i])
        if self.record is not None:
            self.record.update(x0=[list(v) for v in self.xprev], noise=[], u_res=[])
        self.w = [-math.log(N)] * N
        self.we = [1 / N] * N
        self.t = 1
        self.nres = 0

    # measurement_equation!  PFtypes.jl:107-120 (PF) / :226-239 (Advanced, Gaussian likelihood)
    def measurement_equation(self, u, y, t, w):
        if any(math.isnan(v) for v in y):                              # any(ismissing, y)  :109
            return
        m = self.m
        for i in range(self.N):
            g = matvec(m.C, self.x[i])
            w[i] += m.logpdf_meas([y[k] - g[k] for k in range(m.ny)])

    def _noise(self, i):                                               # rand!(rng, d, noise)  PFtypes.jl:135,153
        if self.inject is not None:
            nz = list(map(float, self.inject["noise"][self.k][i]))
        else:
            nz = lower_times(self.m.L1, normals(self.seed, self.epoch, ST_DYN, self.t, i, self.m.nx))
        if self.record is not None:
            while len(self.record["noise"]) <= self.k:
                self.record["noise"].append([])
            self.record["noise"][self.k].append(list(nz))
        return nz

    # propagate_particles!  PFtypes.jl:122-139, :242-289, ext/...DistributionsExt.jl:83-93
    def propagate(self, u, t, use_j, with_noise):
        m = self.m
        for i in range(self.N):
            src = self.j[i] - 1 if use_j else i
            fx = m.dynamics(self.xprev[src], u, t)
            if with_noise:
                nz = self._noise(i)
                self.x[i] = [fx[k] + nz[k] for k in range(m.nx)]
            else:
                self.x[i] = fx

    def shouldresample(self):                                          # resample.jl:5-10
        if self.thr == 1:
            return True
        return effective_particles(self.we) < self.N * self.thr

    def _resample(self, we):                                           # resample.jl:12-15 + our RNG streams
        N = self.N
        if self.inject is not None:
            us = list(map(float, self.inject["u_res"][self.k]))
        elif self.resampling == 2:
            us = [uniform53(self.seed, self.epoch, ST_RESID, self.t, i) for i in range(N)]
        elif self.resampling == 1:
            us = [uniform53(self.seed, self.epoch, ST_STRAT, self.t, i) for i in range(N)]
        else:
            us = [uniform53(self.seed, self.epoch, ST_RESAMPLE, self.t, 0)]
        if self.record is not None:
            while len(self.record["u_res"]) <= self.k:
                self.record["u_res"].append([])
            self.record["u_res"][self.k] = list(us)
        if self.resampling == 2:
            resample_residual(we, self.j, self.bins, us)
        elif self.resampling == 1:
            resample_stratified(we, self.j, self.bins, us)
        else:
            resample_systematic(we, self.j, self.bins, us[0])
        self.nres += 1

    def reset_weights(self):                                           # utils.jl:73-78
        N = self.N
        self.w = [math.log(1 / N)] * N
        self.we = [1 / N] * N

    def predict(self, u, t):                                           # filtering.jl:140-153
        if self.shouldresample():
            self._resample(self.we)
            self.propagate(u, t, True, True)
            self.reset_weights()
        else:
            self.j = list(range(1, self.N + 1))
            self.propagate(u, t, False, True)
        self.xprev = [list(v) for v in self.x]
        self.t += 1
        self._end_predict()

    def _end_predict(self):
        if self.record is not None:
            for key in ("noise", "u_res"):
                while len(self.record[key]) <= self.k:
                    self.record[key].append([])
        self.k += 1

    def correct(self, u, y, t):                                        # filtering.jl:164-174
        if not self.aux:
            self.measurement_equation(u, y, t, self.w)
        return logsumexp(self.w, self.we)

    def predict_aux(self, u, y1, t):                                   # filtering.jl:195-217 / :219-234
        N = self.N
        self.propagate(u, t, False, False)                             # :199
        lam = self.we                                                  # :200  λ = s.we (alias)
        for i in range(N):
            lam[i] = 0.0                                               # :201
        self.measurement_equation(u, y1, t, lam)                       # :202
        for i in range(N):
            self.w[i] += lam[i]                                        # :203
        expnormalize1(self.w)                                          # :204
        self._resample(self.w)                                         # :205
        if self.kind == AUX_ADVANCED:
            self.reset_weights()                                       # :228 (rebinds we: λ is gone, as in the alias)
            self.propagate(u, t, True, True)                           # :230
        else:
            for i in range(N):                                         # :207 permute_with_buffer!  utils.jl:81-86
                self.xprev[i] = list(self.x[self.j[i] - 1])
            self.x = [list(v) for v in self.xprev]
            for i in range(N):                                         # :208 add_noise!  PFtypes.jl:146-157
                nz = self._noise(i)
                self.x[i] = [self.x[i][k] + nz[k] for k in range(self.m.nx)]
            lN = math.log(N)
            for i in range(N):
                self.w[i] = lam[i] - lN                                # :210-213
        self.t += 1                                                    # :215
        self.xprev = [list(v) for v in self.x]                         # :216
        self._end_predict()

    def update(self, u, y, t, y1=None):                                # filtering.jl:181-191
        ll = self.correct(u, y, t)
        if self.aux:
            self.predict_aux(u, y1, t)
        else:
            self.predict(u, t)
        return ll

    def forward_trajectory(self, u, y, epoch=0):                       # filtering.jl:343-365 / :367-384
        self.reset(epoch)
        T = len(y)
        out = dict(x=[], w=[], we=[], ll_steps=[], resampled=[])
        ll = 0.0
        for t in range(1, T + 1):
            ti = (t - 1) * self.Ts
            lli = self.correct(u[t - 1], y[t - 1], ti)
            ll += lli
            out["ll_steps"].append(lli)
            out["x"].append([list(v) for v in self.x])
            out["w"].append(list(self.w))
            out["we"].append(list(self.we))
            n0 = self.nres
            if self.aux:
                if t < T:
                    self.predict_aux(u[t - 1], y[t], ti)
            else:
                self.predict(u[t - 1], ti)
            out["resampled"].append(self.nres - n0)
        out["ll"] = ll
        return out

    def loglik(self, u, y, epoch=0):                                   # smoothing.jl:227-230 / :232-236
        self.reset(epoch)
        T = len(y)
        ll = 0.0
        res = []
        if not self.aux:
            for t in range(T):
                n0 = self.nres
                ll += self.update(u[t], y[t], self.t * self.Ts)        # t = index(pf)*Ts  filtering.jl:181,238
                res.append(self.nres - n0)
        else:
            for t in range(1, T):                                      # sum over 1:length(u)-1
                n0 = self.nres
                ll += self.update(u[t - 1], y[t - 1], (t - 1) * self.Ts, y1=y[t])
                res.append(self.nres - n0)
            keep = self.kind                                           # pf.pf(u[end], y[end], p, (length(u)-1)*Ts)
            self.kind = PF if keep == AUX else ADVANCED
            n0 = self.nres
            ll += self.update(u[T - 1], y[T - 1], (T - 1) * self.Ts)
            res.append(self.nres - n0)
            self.kind = keep
        return dict(ll=ll, resampled=res)
