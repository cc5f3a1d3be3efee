"""rbpf_ref.py — TEST INFRASTRUCTURE: CPU restatement of the reference's Rao-Blackwellized particle filter.

Follows src/rbpf.jl (paths relative to the reference root) on its own data structure — a particle is the triple
(xn, xl, R) of `RBParticle` (rbpf.jl:1-5), every particle runs its own Kalman measurement update and Riccati step — and
NOT the composite-state device code the product generates (lowlevelparticlefilters.jl_b200/rbpf.py), so that the two
check each other.  The generic particle-filter skeleton (`predict!` = shouldresample / resample / propagate / copy,
`correct!` = weights + logsumexp!, `forward_trajectory`, `loglik`) is inherited from the pure-Python twin oracle/pyref.py;
this file supplies what rbpf.jl overrides: reset! (:136-150), predict!'s particle loop (:163-233) and correct! (:236-283),
plus the Kalman correct! they call (filtering.jl:100-133).

Ours (not the reference's): the counter-based RNG streams.  A particle of the device filter is the vector
[xn; xl; lower triangle of R] of length NX, and the device draws NX normals per particle and step of which the first nxn
are used — so does this file (normals(..., NX)[:nxn]).

Pure-Python loops: small N only.  Only tests/ may import this module.
"""
import math

from . import pyref as P


def _matmul(A, B):
    return [[sum((A[i][k] * B[k][j] for k in range(1, len(B))), A[i][0] * B[0][j]) for j in range(len(B[0]))]
            for i in range(len(A))]


def _matvec(A, v):
    return P.matvec(A, v) if len(v) else [0.0] * len(A)


def _T(A):
    return [list(r) for r in zip(*A)]


def _add(A, B):
    return [[a + b for a, b in zip(ra, rb)] for ra, rb in zip(A, B)]


def _sub(A, B):
    return [[a - b for a, b in zip(ra, rb)] for ra, rb in zip(A, B)]


def _symmetrize(A):                                                   # filtering.jl:73-76  0.5 .* (x .+ x')
    n = len(A)
    return [[0.5 * (A[i][j] + A[j][i]) for j in range(n)] for i in range(n)]


def _is_zero(A):
    return A is None or all(v == 0 for r in A for v in r)


def _right_divide_chol(M, L):
    """M / cholesky(S) with S = L L': solve X S = M row by row (forward then backward substitution)"""
    n = len(L)
    out = []
    for row in M:
        yv = [0.0] * n
        for i in range(n):                                            # L y = row'
            v = row[i]
            for k in range(i):
                v -= L[i][k] * yv[k]
            yv[i] = v / L[i][i]
        xv = [0.0] * n
        for i in reversed(range(n)):                                  # L' x = y
            v = yv[i]
            for k in range(i + 1, n):
                v -= L[k][i] * xv[k]
            xv[i] = v / L[i][i]
        out.append(xv)
    return out


def _logpdf_chol(L, e):
    """extended_logpdf(SimpleMvNormal(PDMat(S, chol)), e)  utils.jl:252-257: -(k log 2pi + logdet S)/2 - invquad/2"""
    n = len(e)
    v = [0.0] * n
    q = 0.0
    for i in range(n):
        acc = e[i]
        for k in range(i):
            acc -= L[i][k] * v[k]
        v[i] = acc / L[i][i]
        q += v[i] * v[i]
    logdet = 0.0
    for i in range(n):
        logdet += math.log(L[i][i])
    logdet *= 2
    return -(n * math.log(2 * math.pi) + logdet) / 2 - q / 2


class _Dims:
    def __init__(self, nx):
        self.nx = nx


class RBPFRef(P.Filter):
    """RBPF(N, kf, dynamics, nl_measurement_model, R1n, d0n; An, ...)  rbpf.jl:113-133.
    kf = dict(A, B, C, R1, R2, mu0, Sigma0) (KalmanFilter(A,B,C,0,R1l,R2,d0l)); fn(xn,u,t), g(xn,u,t) Python callables;
    An a matrix or None; d0n = (mu0n, Sigma0n)."""

    def __init__(self, N, kf, fn, g, R1n, d0n, An=None, resample_threshold=0.1, Ts=1.0, seed=0, inject=None, record=None):
        """inject / record as in pyref.Filter: dict(x0 = initial nonlinear states [N][nxn], noise = [K][N][nxn] the draws
        rand(rng, R1n) of every predict!, u_res = [K][...] the rand() of resample) — recorded from the reference itself by
        julia/dump_golden.jl (`rbpf` section) and consumed in place of the counter-based streams."""
        self.kf, self.fn, self.g = kf, fn, g
        self.R1n, self.An = [list(map(float, r)) for r in R1n], (None if An is None else [list(map(float, r)) for r in An])
        self.mu0n, self.Sigma0n = list(map(float, d0n[0])), [list(map(float, r)) for r in d0n[1]]
        self.nxn, self.nxl = len(self.mu0n), len(kf["mu0"])
        self.NX = self.nxn + self.nxl + self.nxl * (self.nxl + 1) // 2
        self.L1n = P.cholesky_lower(self.R1n) if any(v != 0 for r in self.R1n for v in r) else [[0.0] * self.nxn for _ in range(self.nxn)]
        self.L0n = P.cholesky_lower(self.Sigma0n) if any(v != 0 for r in self.Sigma0n for v in r) else [[0.0] * self.nxn for _ in range(self.nxn)]
        self.L2 = P.cholesky_lower(kf["R2"])
        super().__init__(_Dims(self.NX), N, kind=P.PF, resampling=0, resample_threshold=resample_threshold, Ts=Ts, seed=seed,
                         inject=inject, record=record)

    # reset!(pf::RBPF)  rbpf.jl:136-150
    def reset(self, epoch=0):
        self.epoch, self.k = epoch, 0
        N = self.N
        for i in range(N):
            if self.inject is not None:
                xn = list(map(float, self.inject["x0"][i]))
            else:
                z = P.normals(self.seed, epoch, P.ST_INIT, 0, i, self.NX)[:self.nxn]
                lz = P.lower_times(self.L0n, z)
                xn = [self.mu0n[r] + lz[r] for r in range(self.nxn)]     # rand(pf.rng, pf.d0n)
            part = (xn, list(map(float, self.kf["mu0"])), [list(map(float, r)) for r in self.kf["Sigma0"]])
            self.x[i] = part
            self.xprev[i] = part
        if self.record is not None:
            self.record.update(x0=[list(p[0]) for p in self.x], noise=[], u_res=[])
        self.w = [-math.log(N)] * N
        self.we = [1 / N] * N
        self.t = 1
        self.nres = 0

    # the particle loop of predict!(pf::RBPF, u, p, t)  rbpf.jl:184-229 (singleR only skips identical recomputation)
    def propagate(self, u, t, use_j, with_noise):
        kf = self.kf
        Al, Bl, R1l = kf["A"], kf["B"], kf["R1"]
        zeroAn = _is_zero(self.An)
        for i in range(self.N):
            xn, xl, R = self.xprev[self.j[i] - 1 if use_j else i]
            if self.inject is not None:
                noise = list(map(float, self.inject["noise"][self.k][i]))
            else:
                noise = P.lower_times(self.L1n, P.normals(self.seed, self.epoch, P.ST_DYN, self.t, i, self.NX)[:self.nxn])
            if self.record is not None:
                while len(self.record["noise"]) <= self.k:
                    self.record["noise"].append([])
                self.record["noise"][self.k].append(list(noise))
            fi = list(self.fn(xn, u, t))
            Axl_l = _matvec(Al, xl)
            Bu = _matvec(Bl, u) if len(u) else [0.0] * self.nxl
            if zeroAn:
                xn1 = [a + b for a, b in zip(fi, noise)]                   # :203
                xl1 = [a + b for a, b in zip(Axl_l, Bu)]                    # :207
                R1 = _add(_matmul(_matmul(Al, R), _T(Al)), R1l)             # :209
            else:
                An = self.An
                Nt = _add(_matmul(_matmul(An, R), _T(An)), self.R1n)        # :217
                Ln = P.cholesky_lower(Nt)
                L = _right_divide_chol(_matmul(_matmul(Al, R), _T(An)), Ln)   # :218  Al*R*An' / Nt
                R1 = _sub(_add(_matmul(_matmul(Al, R), _T(Al)), R1l), _matmul(_matmul(L, Nt), _T(L)))   # :219
                Axl = _matvec(An, xl)
                z = [a + b for a, b in zip(Axl, noise)]                     # :222
                xn1 = [a + b for a, b in zip(fi, z)]                        # :223
                d = [a - b for a, b in zip(z, Axl)]
                Ld = _matvec(L, d)
                xl1 = [(a + b) + c for a, b, c in zip(Axl_l, Bu, Ld)]       # :225
            self.x[i] = (xn1, xl1, R1)

    # correct!(pf::RBPF, u, y, p, t)  rbpf.jl:236-283 with the Kalman correct! of filtering.jl:100-133
    def measurement_equation(self, u, y, t, w):
        kf = self.kf
        Cm = kf["C"]
        zeroC = _is_zero(Cm)
        ny = len(y)
        for i in range(self.N):
            xn, xl, R = self.x[i]
            yn = list(self.g(xn, u, t))
            if not zeroC:
                ym = [y[k] - yn[k] for k in range(ny)]                      # correct!(kf, u, y-yn, p, t)  :263
                Cx = _matvec(Cm, xl)
                e = [ym[k] - Cx[k] for k in range(ny)]                      # filtering.jl:102
                S = _add(_symmetrize(_matmul(_matmul(Cm, R), _T(Cm))), kf["R2"])   # :120-121
                Ls = P.cholesky_lower(S)
                K = _right_divide_chol(_matmul(R, _T(Cm)), Ls)              # :124
                Ke = _matvec(K, e)
                xl2 = [xl[k] + Ke[k] for k in range(self.nxl)]              # :125
                KC = _matmul(K, Cm)
                ImKC = [[(1.0 if r == c else 0.0) - KC[r][c] for c in range(self.nxl)] for r in range(self.nxl)]
                R2 = _symmetrize(_matmul(ImKC, R))                          # :126
                w[i] += _logpdf_chol(Ls, e)                                 # :128, rbpf.jl:271
                self.x[i] = (xn, xl2, R2)
            else:
                yl = _matvec(Cm, xl) if Cm is not None else [0.0] * ny
                yh = [yn[k] + yl[k] for k in range(ny)]
                w[i] += _logpdf_chol(self.L2, [y[k] - yh[k] for k in range(ny)])   # rbpf.jl:274-275
        self.xprev = list(self.x)                                           # :281

    def correct(self, u, y, t):
        self.measurement_equation(u, y, t, self.w)
        return P.logsumexp(self.w, self.we)

    def predict(self, u, t):                                                # rbpf.jl:163-233 (same skeleton as filtering.jl:140-153)
        if self.shouldresample():
            self._resample(self.we)
            self.reset_weights()
            self.propagate(u, t, True, True)
        else:
            self.j = list(range(1, self.N + 1))
            self.propagate(u, t, False, True)
        self.xprev = list(self.x)
        self.t += 1
        self._end_predict()

    # --- flat views used by the tests -------------------------------------------------------------------------------
    def flat(self, part):
        xn, xl, R = part
        return list(xn) + list(xl) + [R[r][c] for r in range(self.nxl) for c in range(r + 1)]

    def particles_flat(self):
        return [self.flat(p) for p in self.x]

    def weighted_mean(self):
        n = self.nxn + self.nxl
        out = [0.0] * n
        for p, we in zip(self.x, self.we):
            f = self.flat(p)
            for k in range(n):
                out[k] += f[k] * we
        return out

    def run(self, u, y, epoch=0):
        """forward_trajectory (filtering.jl:343-365) keeping ll, per-step ll, resample flags and the weighted mean after
        each correct!"""
        self.reset(epoch)
        out = dict(ll_steps=[], resampled=[], xhat=[])
        ll = 0.0
        for t in range(1, len(y) + 1):
            ti = (t - 1) * self.Ts
            lli = self.correct(u[t - 1], y[t - 1], ti)
            ll += lli
            out["ll_steps"].append(lli)
            out["xhat"].append(self.weighted_mean())
            n0 = self.nres
            self.predict(u[t - 1], ti)
            out["resampled"].append(self.nres - n0)
        out["ll"] = ll
        return out


def kalman_loglik(A, B, C, R1, R2, mu0, Sigma0, u, y):
    """forward_trajectory(kf, u, y).ll  (filtering.jl:52-87,100-133; kalman.jl:159-164) — the closed form the reference's
    own RBPF tests compare with (test_rbpf.jl:111,141)."""
    x, R = list(map(float, mu0)), [list(map(float, r)) for r in Sigma0]
    ll = 0.0
    ny = len(y[0])
    for t in range(len(y)):
        Cx = _matvec(C, x)
        e = [y[t][k] - Cx[k] for k in range(ny)]
        S = _add(_symmetrize(_matmul(_matmul(C, R), _T(C))), R2)
        Ls = P.cholesky_lower(S)
        K = _right_divide_chol(_matmul(R, _T(C)), Ls)
        Ke = _matvec(K, e)
        x = [x[k] + Ke[k] for k in range(len(x))]
        KC = _matmul(K, C)
        n = len(x)
        R = _symmetrize(_matmul([[(1.0 if r == c else 0.0) - KC[r][c] for c in range(n)] for r in range(n)], R))
        ll += _logpdf_chol(Ls, e)
        Ax = _matvec(A, x)
        Bu = _matvec(B, u[t]) if len(u[t]) else [0.0] * n
        x = [a + b for a, b in zip(Ax, Bu)]
        R = _add(_symmetrize(_matmul(_matmul(A, R), _T(A))), R1)
    return ll
