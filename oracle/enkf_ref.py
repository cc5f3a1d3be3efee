"""enkf_ref.py — TEST INFRASTRUCTURE: CPU restatement of the reference's Ensemble Kalman filter (src/enkf.jl).

Stochastic EnKF with perturbed observations, written from enkf.jl (reset! :205-224, predict! :228-272, correct! :281-356,
_ensemble_mean / _ensemble_cov :146-170) and forward_trajectory(kf::AbstractKalmanFilter) (filtering.jl:282-325), in the
reference's evaluation order: sequential sums over the ensemble.  Ours (not the reference's): the counter-
based RNG — initial ensemble stream 0, process noise stream 1, observation perturbations stream 8, keyed by
(seed, epoch; step = enkf.t, member).  Pure-Python loops: small N only.  Only tests/ may import this module.
"""
import math

from . import pyref as P
from .rbpf_ref import _logpdf_chol, _right_divide_chol

ST_ENKF_OBS = 8


class EnKFRef:
    def __init__(self, dynamics, C, R1, R2, mu0, Sigma0, N, Ts=1.0, inflation=1.0, seed=0, inject=None, record=None):
        """dynamics(x, u, t) -> list ; measurement h(x) = C x.
        inject / record: dict(z0 = [N][nx], zdyn = [K][N][nx], zobs = [K][N][ny]) — the STANDARD NORMALS behind the initial
        ensemble, the process noise of the K-th predict! and the observation perturbations of the K-th correct! (what a recording
        RNG sees inside the reference: julia/dump_golden.jl, `enkf` section), consumed in place of the counter-based streams."""
        self.inject, self.record = inject, record
        self.kp = self.kc = 0
        self.f, self.C = dynamics, [list(map(float, r)) for r in C]
        self.R2 = [list(map(float, r)) for r in R2]
        self.L1, self.L2, self.L0 = P.cholesky_lower(R1), P.cholesky_lower(R2), P.cholesky_lower(Sigma0)
        self.mu0 = list(map(float, mu0))
        self.N, self.Ts, self.inflation, self.seed = N, Ts, inflation, seed
        self.nx, self.ny = len(self.mu0), len(self.C)
        self.reset(0)

    def _stats(self):                                                   # _update_ensemble_stats!  enkf.jl:172-176
        N, nx = self.N, self.nx
        xb = list(self.X[0])
        for i in range(1, N):                                           # _ensemble_mean  :146-155
            xb = [a + b for a, b in zip(xb, self.X[i])]
        xb = [a / N for a in xb]
        R = [[0.0] * nx for _ in range(nx)]
        for i in range(N):                                              # _ensemble_cov  :157-168
            d = [a - b for a, b in zip(self.X[i], xb)]
            for r in range(nx):
                for c in range(nx):
                    R[r][c] += d[r] * d[c]
        self.x = xb
        self.R = [[v / (N - 1) for v in row] for row in R]

    def reset(self, epoch=0):                                           # reset!  enkf.jl:205-224
        self.epoch = epoch
        self.kp = self.kc = 0
        self.X = []
        if self.record is not None:
            self.record.update(z0=[], zdyn=[], zobs=[])
        for i in range(self.N):
            z = self._z("z0", None, i, P.ST_INIT, 0, self.nx)
            lz = P.lower_times(self.L0, z)
            self.X.append([self.mu0[r] + lz[r] for r in range(self.nx)])
        self.t = 0
        self._stats()

    def _z(self, key, k, i, stream, step, n):
        if self.inject is not None:
            z = list(map(float, self.inject[key][i] if k is None else self.inject[key][k][i]))
        else:
            z = P.normals(self.seed, self.epoch, stream, step, i, n)
        if self.record is not None:
            tab = self.record[key]
            if k is not None:
                while len(tab) <= k:
                    tab.append([])
                tab = tab[k]
            tab.append(list(z))
        return z

    def predict(self, u, t):                                            # predict!  enkf.jl:228-272
        N, nx = self.N, self.nx
        for i in range(N):
            nz = P.lower_times(self.L1, self._z("zdyn", self.kp, i, P.ST_DYN, self.t, nx))
            fx = self.f(self.X[i], u, t)
            self.X[i] = [fx[r] + nz[r] for r in range(nx)]              # :256
        if self.inflation > 1.0:                                        # :261-266
            xb = list(self.X[0])
            for i in range(1, N):
                xb = [a + b for a, b in zip(xb, self.X[i])]
            xb = [a / N for a in xb]
            for i in range(N):
                self.X[i] = [xb[r] + self.inflation * (self.X[i][r] - xb[r]) for r in range(nx)]
        self.t += 1
        self.kp += 1
        self._stats()

    def correct(self, u, y, t):                                         # correct!  enkf.jl:281-356
        N, nx, ny = self.N, self.nx, self.ny
        Y = [P.matvec(self.C, x) for x in self.X]                       # :298-305
        xb = [sum(x[r] for x in self.X) / N for r in range(nx)]         # mean(X)  :308
        yb = [sum(yv[a] for yv in Y) / N for a in range(ny)]            # :309
        Xa = [[x[r] - xb[r] for r in range(nx)] for x in self.X]        # :312-314
        Ya = [[yv[a] - yb[a] for a in range(ny)] for yv in Y]           # :315
        S = [[sum(Ya[i][a] * Ya[i][b] for i in range(N)) / (N - 1) + self.R2[a][b] for b in range(ny)] for a in range(ny)]   # :318
        S = [[0.5 * (S[a][b] + S[b][a]) for b in range(ny)] for a in range(ny)]                                          # :319
        Ls = P.cholesky_lower(S)                                        # :322
        Rxy = [[sum(Xa[i][r] * Ya[i][a] for i in range(N)) / (N - 1) for a in range(ny)] for r in range(nx)]            # :329-330
        K = _right_divide_chol(Rxy, Ls)                                 # :331
        e = [y[a] - yb[a] for a in range(ny)]                           # :334
        for i in range(N):                                              # :340-349
            eps = P.lower_times(self.L2, self._z("zobs", self.kc, i, ST_ENKF_OBS, self.t, ny))
            d = [(y[a] + eps[a]) - Y[i][a] for a in range(ny)]
            Kd = P.matvec(K, d)
            self.X[i] = [self.X[i][r] + Kd[r] for r in range(nx)]
        ll = _logpdf_chol(Ls, e)                                        # :352
        self.kc += 1
        self._stats()
        return dict(ll=ll, e=e, S=S, K=K)

    def forward_trajectory(self, u, y, epoch=0):                        # filtering.jl:282-325
        self.reset(epoch)
        out = dict(x=[], R=[], xt=[], Rt=[], e=[], ll_steps=[], S=[], K=[])
        ll = 0.0
        for k in range(len(y)):
            ti = k * self.Ts
            out["x"].append(list(self.x)); out["R"].append([list(r) for r in self.R])
            r = self.correct(u[k], y[k], ti)
            ll += r["ll"]
            out["ll_steps"].append(r["ll"]); out["e"].append(r["e"]); out["S"].append(r["S"]); out["K"].append(r["K"])
            out["xt"].append(list(self.x)); out["Rt"].append([list(r2) for r2 in self.R])
            self.predict(u[k], ti)
        out["ll"] = ll
        return out
