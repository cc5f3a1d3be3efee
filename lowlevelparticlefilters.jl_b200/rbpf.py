"""Rao-Blackwellized particle filter on the device — host side (reference src/rbpf.jl).

    xn+ = fn(xn,u,p,t) + An xl + wn,  wn ~ N(0, R1n)
    xl+ = A xl + B u + wl,            wl ~ N(0, R1l)
    y   = g(xn,u,p,t) + C xl + e,     e  ~ N(0, R2)                                   (rbpf.jl:88-93)

The reference plugs RBPF into the generic particle-filter verbs by overriding reset!/predict!/correct! on a particle type
that carries a Kalman filter (RBParticle, rbpf.jl:1-5).  Here the same plug point is the user-model boundary of
the C-ABI: `RBPF(...)` generates CUDA source for `llpf_create_user` (include/llpf.h, LLPF_USER_STATE_HOOKS) in which a
particle is the vector [xn; xl; lower triangle of R]; the arithmetic of rbpf.jl:163-283 lives in csrc/llpf_rbpf.cuh and is
inlined into the fused sweep of the f64 engine.  Every verb (reset, correct, predict, update, forward_trajectory, loglik,
accessors) is the generic one of filters.py.  `fn` and `g` are given as device snippets (closures cannot cross a C-ABI).

Restrictions (reported, not silently ignored): constant matrices A, B, C, An, R1l (the reference also accepts functions
of (x,u,p,t)); D = 0; nxn + nxl + nxl(nxl+1)/2 <= 8 and ny <= 8 (the f64 engine's particle size).
"""
from dataclasses import dataclass

import numpy as np

from . import filters as F


@dataclass
class KalmanFilter:
    """KalmanFilter(A, B, C, D, R1, R2, d0) (kalman.jl:75-86): the linear sub-model handed to RBPF; not a filter here."""
    A: np.ndarray
    B: np.ndarray
    C: np.ndarray
    D: object
    R1: np.ndarray
    R2: np.ndarray
    d0: F.MvNormal


@dataclass
class RBMeasurementModel:
    """RBMeasurementModel(measurement, R2, ny) (rbpf.jl:40-55).  `measurement`: the BODY of a device function that writes
    g(xn,u,p,t) into `double yn[NY]`; in scope: `const double (&xn)[NXN]`, `const double* u`, `const double* p`, `double t`."""
    measurement: str
    R2: np.ndarray
    ny: int


def _lit(v):
    return float(v).hex()


def _mat(M, rows, cols):
    """nested brace initialiser of a rows x max(cols,1) matrix (zero-size dimensions get one dummy column)"""
    A = np.zeros((rows, max(cols, 1)))
    if cols > 0 and M is not None:
        A[:, :cols] = np.asarray(M, dtype=np.float64).reshape(rows, cols)
    return "{" + ", ".join("{" + ", ".join(_lit(v) for v in r) + "}" for r in A) + "}"


def rbpf_source(nxn, nxl, ny, nu, A, B, C, An, R1l, R1n, R2, fn_body, g_body):
    """The llpf_user source of an RBPF (see module docstring); also used by the offline NVRTC compile test."""
    nx = nxn + nxl + nxl * (nxl + 1) // 2
    zero_an = An is None or not np.any(np.asarray(An, dtype=np.float64))
    zero_c = C is None or not np.any(np.asarray(C, dtype=np.float64))
    consts = ", ".join([_mat(A, nxl, nxl), _mat(B, nxl, nu), _mat(C, ny, nxl), _mat(An, nxn, nxl), _mat(R1l, nxl, nxl),
                        _mat(R1n, nxn, nxn), _mat(R2, ny, ny), str(int(zero_an)), str(int(zero_c))])
    return f"""// LLPF_USER_STATE_HOOKS : Rao-Blackwellized particle filter (rbpf.jl), particle = [xn({nxn}); xl({nxl}); tril(R)]
#include "llpf_rbpf.cuh"
namespace llpf_user {{
using RB = llpf_rbpf::Consts<{nxn}, {nxl}, {ny}, {nu}>;
__device__ __forceinline__ RB rb_consts() {{ const RB k = {{{consts}}}; return k; }}
__device__ __forceinline__ void rb_fn(const double (&xn)[{nxn}], const double* u, const double* p, double t, double (&fn)[{nxn}]) {{
{fn_body}
}}
__device__ __forceinline__ void rb_g(const double (&xn)[{nxn}], const double* u, const double* p, double t, double (&yn)[{ny}]) {{
{g_body}
}}
__device__ __forceinline__ void rb_xn(const double (&x)[{nx}], double (&xn)[{nxn}]) {{
#pragma unroll
  for (int r = 0; r < {nxn}; ++r) xn[r] = x[r];
}}
template <> __device__ void dynamics<{nx}>(double (&x)[{nx}], const double* u, const double* p, double t) {{
  const RB k = rb_consts();
  double xn[{nxn}], fn[{nxn}];
  rb_xn(x, xn);
  rb_fn(xn, u, p, t, fn);
  llpf_rbpf::predict_mean<{nxn}, {nxl}, {ny}, {nu}, {nx}>(x, fn, u, k);
}}
template <> __device__ void add_noise<{nx}>(double (&x)[{nx}], const double (&xprev)[{nx}], const double (&nz)[{nx}],
                                           const double* u, const double* p, double t) {{
  const RB k = rb_consts();
  llpf_rbpf::add_noise<{nxn}, {nxl}, {ny}, {nu}, {nx}>(x, xprev, nz, k);
}}
template <> __device__ double loglik<{nx}>(const double (&x)[{nx}], const double* u, const double* y, const double* p, double t) {{
  const RB k = rb_consts();
  double xn[{nxn}], yn[{ny}];
  rb_xn(x, xn);
  rb_g(xn, u, p, t, yn);
  return llpf_rbpf::loglik<{nxn}, {nxl}, {ny}, {nu}, {nx}>(x, yn, y, k);
}}
template <> __device__ void correct_state<{nx}>(double (&x)[{nx}], const double* u, const double* y, const double* p, double t) {{
  const RB k = rb_consts();
  double xn[{nxn}], yn[{ny}];
  rb_xn(x, xn);
  rb_g(xn, u, p, t, yn);
  llpf_rbpf::correct_state<{nxn}, {nxl}, {ny}, {nu}, {nx}>(x, yn, y, k);
}}
}}  // namespace llpf_user
"""


class RBPF(F.AbstractParticleFilter):
    """RBPF(N, kf, dynamics, nl_measurement_model, R1n, d0n; An, nu, Ts=1.0, p, resample_threshold=0.1)  (rbpf.jl:113-133).

    kf: KalmanFilter (linear sub-model); dynamics: BODY of a device function that writes fn(xn,u,p,t) into `double fn[NXN]`
    (in scope: xn, u, p, t); nl_measurement_model: RBMeasurementModel; R1n: covariance of the nonlinear state's noise;
    d0n: MvNormal of the initial nonlinear state (a zero covariance gives a deterministic start); An: matrix or None.
    `seed` replaces `rng`.  particles(pf) are the composite vectors; rb_particles(pf) splits them."""
    _filter_code = F.FILTER_PF

    def __init__(self, N, kf, dynamics, nl_measurement_model, R1n, d0n, *, An=None, nu=0, Ts=1.0, p=None,
                 resample_threshold=0.1, seed=0, scan_mode="fast", device=0, names=None, **_ignored):
        if not isinstance(kf, KalmanFilter):
            raise TypeError("kf must be an llpf_b200.KalmanFilter(A, B, C, D, R1, R2, d0) descriptor")
        if any(callable(m) for m in (kf.A, kf.B, kf.C, kf.R1, An)):
            raise TypeError("RBPF on the device takes constant matrices (functions of (x,u,p,t) are not supported)")
        if kf.D is not None and np.any(np.asarray(kf.D, dtype=np.float64)):
            raise ValueError("RBPF on the device supports D = 0 only")
        mm = nl_measurement_model
        nxn, nxl, ny, nu = len(d0n), len(kf.d0), int(mm.ny), int(nu)
        nx = nxn + nxl + nxl * (nxl + 1) // 2
        if nx > 8:
            raise ValueError(f"RBPF particle [xn; xl; tril(R)] has {nx} components; the f64 engine carries at most 8")
        self.kf, self.dynamics, self.nl_measurement_model = kf, dynamics, mm
        self.measurement = mm
        self.R1n, self.d0n, self.An = np.atleast_2d(np.asarray(R1n, dtype=np.float64)), d0n, An
        self.nxn, self.nxl = nxn, nxl
        src = rbpf_source(nxn, nxl, ny, nu, kf.A, kf.B, kf.C, An, kf.R1, self.R1n, mm.R2, dynamics, mm.measurement)
        # composite densities: only xn is random (rbpf.jl:136-150, :203,:222); the other components start at kf.d0 and
        # carry no additive noise (zero rows: accepted as positive SEMI-definite by llpf_create_user)
        R1c = np.zeros((nx, nx)); R1c[:nxn, :nxn] = self.R1n
        mu0 = np.zeros(nx); S0 = np.zeros((nx, nx))
        mu0[:nxn] = d0n.mu; S0[:nxn, :nxn] = d0n.Sigma
        mu0[nxn:nxn + nxl] = kf.d0.mu
        R0 = np.atleast_2d(np.asarray(kf.d0.Sigma, dtype=np.float64))
        mu0[nxn + nxl:] = [R0[r, c] for r in range(nxl) for c in range(r + 1)]
        self.dynamics_density = F.MvNormal(np.zeros(nx), R1c)
        self.initial_density = F.MvNormal(mu0, S0)
        self.measurement_density = F.MvNormal(np.zeros(ny), np.atleast_2d(np.asarray(mm.R2, dtype=np.float64)))
        model = F._UserModelBuffers(None, None, ny, R1c, self.initial_density, source=src, nu=nu)
        self._create(N, model, resample_threshold, F.ResampleSystematic, Ts, seed, scan_mode, device, p=p)


def rb_particles(pf, x=None):
    """(xn [N][nxn], xl [N][nxl], R [N][nxl][nxl]) of the current particles (or of an array of composite particles)"""
    X = F.particles(pf) if x is None else np.asarray(x)
    nxn, nxl = pf.nxn, pf.nxl
    xn, xl = X[..., :nxn], X[..., nxn:nxn + nxl]
    R = np.zeros(X.shape[:-1] + (nxl, nxl))
    k = nxn + nxl
    for r in range(nxl):
        for c in range(r + 1):
            R[..., r, c] = R[..., c, r] = X[..., k]
            k += 1
    return xn, xl, R
