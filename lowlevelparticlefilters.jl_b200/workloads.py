"""Synthetic workloads of BASELINE.json / SURVEY.md §8d (shapes, distributions, seeds).

config 1: examples/example_lineargaussian.jl:10-29  (nx=2 in the file; BASELINE says 4-state: both)
config 2: ParticleFilter, 4-state LG, N=2^20, T=1000, Float64
config 3: AdvancedParticleFilter, quadtank RK4 (examples/example_quadtank.jl), N=2^18, T=2000
config 4: AuxiliaryParticleFilter, 4-state LG, N=2^22
config 5: ParticleFilter, 64-state LG (test/test_large.jl regime), N=2^20, T=500, Float32 particles
"""
from dataclasses import dataclass

import numpy as np

from . import filters as F


@dataclass
class LGSpec:
    nx: int
    nu: int
    ny: int
    A: np.ndarray
    B: np.ndarray
    C: np.ndarray
    R1: np.ndarray
    R2: np.ndarray
    mu0: np.ndarray
    Sigma0: np.ndarray
    dtype: object = np.float64     # particle element type (eltype of the initial density, PFtypes.jl:66)

    def _parts(self):
        return (F.LinearDynamics(self.A, self.B), F.LinearMeasurement(self.C), F.MvNormal(np.zeros(self.nx), self.R1),
                F.MvNormal(np.zeros(self.ny), self.R2), F.MvNormal(self.mu0, self.Sigma0, dtype=self.dtype))

    def particle_filter(self, N, **kw):
        dyn, meas, df, dg, d0 = self._parts()
        return F.ParticleFilter(N, dyn, meas, df, dg, d0, **kw)

    def aux_filter(self, N, **kw):
        dyn, meas, df, dg, d0 = self._parts()
        return F.AuxiliaryParticleFilter(N, dyn, meas, df, dg, d0, **kw)

    def advanced_filter(self, N, **kw):
        dyn, meas, df, dg, d0 = self._parts()
        return F.AdvancedParticleFilter(N, dyn, meas, F.GaussianLikelihood(self.C, self.R2), df, d0, **kw)


def lg_spec(nx=4, nu=2, ny=2, seed=0, r1=1.0, r2=1.0):
    """A = Tr*diag(LinRange(0.5,0.95,nx))/Tr, B,C,Tr ~ randn; df=dg=N(0,I); d0=N(randn, 2^2 I)
    (examples/example_lineargaussian.jl:15-23)"""
    rng = np.random.default_rng(seed)
    Tr = rng.standard_normal((nx, nx))
    A = Tr @ np.diag(np.linspace(0.5, 0.95, nx)) @ np.linalg.inv(Tr)
    B = rng.standard_normal((nx, nu))
    C_ = rng.standard_normal((ny, nx))
    mu0 = rng.standard_normal(nx)
    return LGSpec(nx, nu, ny, A, B, C_, r1 * np.eye(nx), r2 * np.eye(ny), mu0, 4.0 * np.eye(nx))


def lg_large_spec(nx=64, nu=2, ny=58, seed=0, dtype=np.float32):
    """BASELINE config 5 — the regime of test/test_large.jl:8-22 moved to a ParticleFilter with Float32 particles:
    A = 0.1*randn(nx,nx), B = randn(nx,nu), C = randn(ny,nx) with ny = 0.9*nx rounded (58 for nx = 64, as the test's
    90/100), R1 = I, R2 = I, d0 = N(0, I)."""
    rng = np.random.default_rng(seed)
    A = 0.1 * rng.standard_normal((nx, nx))
    B = rng.standard_normal((nx, nu))
    C_ = rng.standard_normal((ny, nx))
    return LGSpec(nx, nu, ny, A, B, C_, np.eye(nx), np.eye(ny), np.zeros(nx), np.eye(nx), dtype=dtype)


@dataclass
class QuadtankSpec:
    """examples/example_quadtank.jl:91-130: p_true, R1 = diag(0.1), R2 = 1e-4 I, d0 = N(x0, R1), y = x[1:2]"""
    p: tuple = (0.5, 1.6, 1.6, 4.9, 0.03, 0.2)
    Ts: float = 1.0
    supersample: int = 2
    t_switch: float = float("inf")
    a1_factor: float = 1.0
    nx: int = 4
    nu: int = 2
    ny: int = 2

    @property
    def C(self):
        return np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0]])

    @property
    def R1(self):
        return np.diag([0.1, 0.1, 0.1, 0.1])

    @property
    def R2(self):
        return (1e-2) ** 2 * np.eye(2)

    @property
    def x0(self):
        return np.array([2.0, 2.0, 3.0, 3.0])

    def dynamics(self):
        return F.QuadtankRK4(self.p, self.Ts, self.supersample, self.t_switch, self.a1_factor)

    def advanced_filter(self, N, **kw):
        return F.AdvancedParticleFilter(N, self.dynamics(), F.LinearMeasurement(self.C),
                                        F.GaussianLikelihood(self.C, self.R2), F.MvNormal(np.zeros(4), self.R1),
                                        F.MvNormal(self.x0, self.R1), **kw)

    def inputs(self, T):
        """u = 0.25*sign(sin(2pi/200 t)) + 0.25 on both pumps  (example_quadtank.jl:37-40)"""
        t = np.arange(T) * self.Ts
        u1 = 0.25 * np.sign(np.sin(2 * np.pi / 200 * t)) + 0.25
        return np.stack([u1, u1], axis=1)


def simulate_lg(spec, u, seed=1):
    """simulate(f,u,p)  src/filtering.jl:462-477 for the LG model with numpy noise (data generation only)."""
    rng = np.random.default_rng(seed)
    T = u.shape[0]
    L1, L2 = np.linalg.cholesky(spec.R1), np.linalg.cholesky(spec.R2)
    x = np.zeros((T, spec.nx))
    y = np.zeros((T, spec.ny))
    x[0] = spec.mu0
    for t in range(T):
        y[t] = spec.C @ x[t] + L2 @ rng.standard_normal(spec.ny)
        if t + 1 < T:
            x[t + 1] = spec.A @ x[t] + spec.B @ u[t] + L1 @ rng.standard_normal(spec.nx)
    return x, y


def simulate_quadtank(spec, u, seed=1):
    """simulate(f,u,p)  src/filtering.jl:462-477 for the quadtank AdvancedParticleFilter (data generation only, numpy):
    x1 = mean(d0) = x0 ; y_t = x_t[1:2] + 0.01 randn (example_quadtank.jl:44) ; x_{t+1} = rk4(quadtank)(x_t,u_t) + N(0,R1)."""
    rng = np.random.default_rng(seed)
    kc, k1, k2, A_, a, gam = spec.p
    g = 9.81

    def rhs(h, uu, t):
        a1 = a * spec.a1_factor if t > spec.t_switch else a
        sq = lambda v: np.sqrt(max(v, 0.0) + 1e-3)  # noqa: E731
        tg = 2 * g
        return np.array([-a1 / A_ * sq(tg * h[0]) + a / A_ * sq(tg * h[2]) + gam * k1 / A_ * uu[0],
                         -a / A_ * sq(tg * h[1]) + a / A_ * sq(tg * h[3]) + gam * k2 / A_ * uu[1],
                         -a / A_ * sq(tg * h[2]) + (1 - gam) * k2 / A_ * uu[1],
                         -a / A_ * sq(tg * h[3]) + (1 - gam) * k1 / A_ * uu[0]])

    T = u.shape[0]
    x = np.zeros((T, 4))
    y = np.zeros((T, 2))
    x[0] = spec.x0
    L1, L2 = np.linalg.cholesky(spec.R1), np.linalg.cholesky(spec.R2)
    h = spec.Ts / spec.supersample
    for t in range(T):
        y[t] = spec.C @ x[t] + L2 @ rng.standard_normal(2)
        if t + 1 < T:
            xx, tt = x[t].copy(), t * spec.Ts
            for _ in range(spec.supersample):
                f1 = rhs(xx, u[t], tt)
                f2 = rhs(xx + h / 2 * f1, u[t], tt + h / 2)
                f3 = rhs(xx + h / 2 * f2, u[t], tt + h / 2)
                f4 = rhs(xx + h * f3, u[t], tt + h)
                xx = xx + h / 6 * (f1 + 2 * f2 + 2 * f3 + f4)
                tt += h
            x[t + 1] = xx + L1 @ rng.standard_normal(4)
    return x, y
