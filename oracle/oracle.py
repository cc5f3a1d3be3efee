"""ctypes wrapper around oracle/liboracle.so — TEST INFRASTRUCTURE (see llpf_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import
this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from llpf_b200._abi import Config, Model  # noqa: E402  (plain struct layouts of include/llpf.h)

LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)


class JuliaRange(C.Structure):
    """orc_range of llpf_oracle.c: Julia's StepRangeLen{Float64,TwicePrecision,TwicePrecision}"""
    _fields_ = [("ref_hi", C.c_double), ("ref_lo", C.c_double), ("step_hi", C.c_double), ("step_lo", C.c_double),
                ("len", C.c_int64), ("offset", C.c_int64), ("rational", C.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "llpf_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        H = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(Config), C.POINTER(Model), C.POINTER(H)]
        L.orc_create.restype = C.c_int
        L.orc_set_model.argtypes = [H, C.POINTER(Model)]
        L.orc_set_model.restype = C.c_int
        L.orc_destroy.argtypes = [H]
        L.orc_destroy.restype = None
        L.orc_reset.argtypes = [H, C.c_uint64]
        L.orc_reset.restype = None
        L.orc_correct.argtypes = [H, dp, dp, C.c_double]
        L.orc_correct.restype = C.c_double
        L.orc_predict.argtypes = [H, dp, C.c_double]
        L.orc_predict.restype = None
        L.orc_predict_aux.argtypes = [H, dp, dp, C.c_double]
        L.orc_predict_aux.restype = None
        L.orc_update.argtypes = [H, dp, dp, dp, C.c_double]
        L.orc_update.restype = C.c_double
        L.orc_shouldresample.argtypes = [H]
        L.orc_shouldresample.restype = C.c_int
        L.orc_weighted_mean.argtypes = [H, dp]
        L.orc_weighted_mean.restype = None
        L.orc_forward_trajectory.argtypes = [H, C.c_int64, dp, dp, C.c_uint64, dp, dp, i32p, dp, dp, dp, dp]
        L.orc_forward_trajectory.restype = C.c_double
        L.orc_loglik.argtypes = [H, C.c_int64, dp, dp, C.c_uint64, dp, dp, i32p]
        L.orc_loglik.restype = C.c_double
        for name in ("orc_particles", "orc_xprev", "orc_weights", "orc_expweights", "orc_bins"):
            getattr(L, name).argtypes = [H]
            getattr(L, name).restype = dp
        L.orc_ancestors.argtypes = [H]
        L.orc_ancestors.restype = ip
        L.orc_num_particles.argtypes = [H]
        L.orc_num_particles.restype = C.c_int64
        L.orc_index.argtypes = [H]
        L.orc_index.restype = C.c_int64
        L.orc_set_state.argtypes = [H, dp, dp, C.c_int64]
        L.orc_set_state.restype = None
        L.orc_simulate.argtypes = [H, C.c_int64, dp, C.c_uint64, dp, dp]
        L.orc_simulate.restype = None
        L.orc_dynamics.argtypes = [H, dp, dp, C.c_double, dp]
        L.orc_dynamics.restype = None
        L.orc_kalman_loglik.argtypes = [C.POINTER(Model), C.c_int64, dp, dp]
        L.orc_kalman_loglik.restype = C.c_double
        L.orc_logsumexp.argtypes = [dp, dp, C.c_int64, dp]
        L.orc_logsumexp.restype = C.c_double
        L.orc_expnormalize1.argtypes = [dp, C.c_int64]
        L.orc_expnormalize1.restype = None
        L.orc_expnormalize2.argtypes = [dp, dp, C.c_int64]
        L.orc_expnormalize2.restype = None
        L.orc_effective_particles.argtypes = [dp, C.c_int64]
        L.orc_effective_particles.restype = C.c_double
        L.orc_resample_systematic.argtypes = [dp, C.c_int64, C.c_double, C.c_int64, ip, dp]
        L.orc_resample_systematic.restype = None
        L.orc_julia_range.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(JuliaRange)]
        L.orc_julia_range.restype = None
        L.orc_julia_range_getindex.argtypes = [C.POINTER(JuliaRange), C.c_int64]
        L.orc_julia_range_getindex.restype = C.c_double
        L.orc_resample_stratified.argtypes = [dp, C.c_int64, dp, C.c_int64, ip, dp]
        L.orc_resample_stratified.restype = None
        L.orc_resample_residual.argtypes = [dp, C.c_int64, dp, C.c_int64, ip, dp]
        L.orc_resample_residual.restype = C.c_int64
        L.orc_set_user_functions.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_user_functions.restype = C.c_int
        L.orc_smooth.argtypes = [H, C.c_int64, C.c_int64, dp, dp, dp, dp, C.c_uint64, dp]
        L.orc_smooth.restype = C.c_int
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_philox4x32_10.restype = None
        L.orc_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, dp]
        L.orc_normals.restype = None
        L.orc_uniform53.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_uniform53.restype = C.c_double
        L.orc_rk4_constant_rhs.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_rk4_constant_rhs.restype = C.c_double
        L.orc_cholesky_lower.argtypes = [dp, C.c_int, dp]
        L.orc_cholesky_lower.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(dp)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


# ---------------------------------------------------------------------------------------------
# stand-alone numerics
# ---------------------------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*[int(v) for v in ctr])
    k = (C.c_uint32 * 2)(*[int(v) for v in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def normals(seed, epoch, stream, t, i, n):
    z = np.zeros(n)
    lib().orc_normals(seed, epoch, stream, t, i, n, _p(z))
    return z


def uniform53(seed, epoch, stream, t, i):
    return lib().orc_uniform53(seed, epoch, stream, t, i)


def logsumexp(w):
    """logsumexp!(w, we): returns (ll, w_normalised, we)."""
    w = _f64(w).copy()
    we = np.empty_like(w)
    mx = C.c_double()
    ll = lib().orc_logsumexp(_p(w), _p(we), w.size, C.byref(mx))
    return ll, w, we


def expnormalize(w, inplace=True):
    w = _f64(w).copy()
    if inplace:
        lib().orc_expnormalize1(_p(w), w.size)
        return w
    we = np.empty_like(w)
    lib().orc_expnormalize2(_p(we), _p(w), w.size)
    return we, w


def effective_particles(we):
    we = _f64(we)
    return lib().orc_effective_particles(_p(we), we.size)


def julia_range(start, step, stop):
    """Julia Base `start:step:stop` (Float64) as restated in llpf_oracle.c; returns (JuliaRange, getindex)."""
    R = JuliaRange()
    lib().orc_julia_range(float(start), float(step), float(stop), C.byref(R))
    return R, (lambda i: lib().orc_julia_range_getindex(C.byref(R), int(i)))


def resample_systematic(we, u01, M=None, j0=None):
    we = _f64(we)
    N = we.size
    M = N if M is None else M
    j = np.arange(1, M + 1, dtype=np.int64) if j0 is None else np.array(j0, dtype=np.int64)
    bins = np.zeros(N)
    lib().orc_resample_systematic(_p(we), N, float(u01), M, j.ctypes.data_as(ip), _p(bins))
    return j, bins


def resample_stratified(we, u01, M=None, j0=None):
    we = _f64(we)
    N = we.size
    M = N if M is None else M
    u01 = _f64(u01)
    assert u01.size == M
    j = np.arange(1, M + 1, dtype=np.int64) if j0 is None else np.array(j0, dtype=np.int64)
    bins = np.zeros(N)
    lib().orc_resample_stratified(_p(we), N, _p(u01), M, j.ctypes.data_as(ip), _p(bins))
    return j, bins


def resample_residual(we, u01, M=None, j0=None, return_bins=False):
    we = _f64(we)
    N = we.size
    M = N if M is None else M
    u01 = _f64(u01)
    j = np.zeros(M, dtype=np.int64) if j0 is None else np.array(j0, dtype=np.int64)
    bins = np.zeros(N)
    lib().orc_resample_residual(_p(we), N, _p(u01), M, j.ctypes.data_as(ip), _p(bins))
    return (j, bins) if return_bins else j


def rk4_constant_rhs(c, x0, Ts, supersample=1):
    return lib().orc_rk4_constant_rhs(c, x0, Ts, supersample)


def cholesky_lower(S):
    S = np.asfortranarray(np.asarray(S, dtype=np.float64))
    n = S.shape[0]
    L = np.zeros((n, n), order="F")
    rc = lib().orc_cholesky_lower(S.ctypes.data_as(dp), n, L.ctypes.data_as(dp))
    if rc:
        raise ValueError("not positive definite")
    return L


# ---------------------------------------------------------------------------------------------
# model / filter objects
# ---------------------------------------------------------------------------------------------
class ModelArrays:
    """Owns column-major copies of the model matrices and the ctypes struct pointing at them."""

    def __init__(self, nx, nu, ny, C_, R1, R2, mu0, Sigma0, A=None, B=None, dynamics=0,
                 dyn_params=None, t_switch=float("inf"), a1_factor=1.0, integ_Ts=1.0, supersample=1):
        F = lambda M, shape: np.asfortranarray(np.asarray(M, dtype=np.float64).reshape(shape))  # noqa: E731
        self.nx, self.nu, self.ny = nx, nu, ny
        self.A = F(A, (nx, nx)) if A is not None else None
        self.B = F(B, (nx, nu)) if (B is not None and nu > 0) else None
        self.C = F(C_, (ny, nx))
        self.R1 = F(R1, (nx, nx))
        self.R2 = F(R2, (ny, ny))
        self.mu0 = _f64(mu0).reshape(nx)
        self.Sigma0 = F(Sigma0, (nx, nx))
        m = Model()
        m.nx, m.nu, m.ny, m.dynamics = nx, nu, ny, dynamics
        null = C.cast(None, dp)
        m.A = self.A.ctypes.data_as(dp) if self.A is not None else null
        m.B = self.B.ctypes.data_as(dp) if self.B is not None else null
        m.C = self.C.ctypes.data_as(dp)
        m.R1 = self.R1.ctypes.data_as(dp)
        m.R2 = self.R2.ctypes.data_as(dp)
        m.mu0 = self.mu0.ctypes.data_as(dp)
        m.Sigma0 = self.Sigma0.ctypes.data_as(dp)
        pars = list(dyn_params) if dyn_params is not None else []
        for k in range(8):
            m.dyn_params[k] = float(pars[k]) if k < len(pars) else 0.0
        m.t_switch, m.a1_factor, m.integ_Ts, m.supersample = t_switch, a1_factor, integ_Ts, supersample
        self.struct = m


class OracleFilter:
    def __init__(self, model, N, filter=0, resampling=0, resample_threshold=0.1, Ts=1.0, seed=0, particle_dtype=np.float64):
        self.model = model
        cfg = Config()
        cfg.particle_dtype = 1 if np.dtype(particle_dtype) == np.dtype(np.float32) else 0
        cfg.N, cfg.filter, cfg.resampling = N, filter, resampling
        cfg.resample_threshold, cfg.Ts, cfg.seed = resample_threshold, Ts, seed
        cfg.scan_mode, cfg.device, cfg.rank, cfg.world = 1, 0, 0, 1
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = lib().orc_create(C.byref(cfg), C.byref(model.struct), C.byref(self.h))
        if rc:
            raise RuntimeError(f"orc_create failed: {rc}")
        self.N, self.nx, self.nu, self.ny = N, model.nx, model.nu, model.ny

    def __del__(self):
        try:
            if self.h:
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _view(self, fn, shape, dtype=np.float64):
        ptr = fn(self.h)
        return np.ctypeslib.as_array(ptr, shape=shape).view(dtype)

    @property
    def particles(self):
        return self._view(lib().orc_particles, (self.N, self.nx))

    @property
    def xprev(self):
        return self._view(lib().orc_xprev, (self.N, self.nx))

    @property
    def weights(self):
        return self._view(lib().orc_weights, (self.N,))

    @property
    def expweights(self):
        return self._view(lib().orc_expweights, (self.N,))

    @property
    def bins(self):
        return self._view(lib().orc_bins, (self.N,))

    @property
    def ancestors(self):
        return np.ctypeslib.as_array(lib().orc_ancestors(self.h), shape=(self.N,))

    @property
    def index(self):
        return lib().orc_index(self.h)

    def reset(self, epoch=0):
        lib().orc_reset(self.h, epoch)

    def set_state(self, x, w, t):
        x, w = _f64(x), _f64(w)
        lib().orc_set_state(self.h, _p(x), _p(w), t)

    def _u(self, u):
        u = _f64(u) if u is not None else np.zeros(0)
        return u if u.size else np.zeros(1)

    def correct(self, u, y, t):
        u, y = self._u(u), _f64(y)
        return lib().orc_correct(self.h, _p(u), _p(y), float(t))

    def predict(self, u, t):
        u = self._u(u)
        lib().orc_predict(self.h, _p(u), float(t))

    def predict_aux(self, u, y1, t):
        u, y1 = self._u(u), _f64(y1)
        lib().orc_predict_aux(self.h, _p(u), _p(y1), float(t))

    def update(self, u, y, t, y1=None):
        u, y = self._u(u), _f64(y)
        y1a = _f64(y1) if y1 is not None else y
        return lib().orc_update(self.h, _p(u), _p(y), _p(y1a), float(t))

    def shouldresample(self):
        return bool(lib().orc_shouldresample(self.h))

    def weighted_mean(self):
        xh = np.zeros(self.nx)
        lib().orc_weighted_mean(self.h, _p(xh))
        return xh

    def forward_trajectory(self, u, y, epoch=0, history=False):
        u, y = _f64(u).reshape(-1, max(self.nu, 1))[:, : self.nu], _f64(y).reshape(-1, self.ny)
        T = y.shape[0]
        uu = np.ascontiguousarray(u) if self.nu else np.zeros((T, 1))
        out = dict(ll_steps=np.zeros(T), ess=np.zeros(T), resampled=np.zeros(T, dtype=np.int32),
                   xhat=np.zeros((T, self.nx)))
        null = C.cast(None, dp)
        xh = wh = weh = None
        if history:
            xh = np.zeros((T, self.N, self.nx))
            wh = np.zeros((T, self.N))
            weh = np.zeros((T, self.N))
        ll = lib().orc_forward_trajectory(
            self.h, T, _p(uu), _p(y), epoch, _p(out["ll_steps"]), _p(out["ess"]),
            out["resampled"].ctypes.data_as(i32p), _p(out["xhat"]),
            _p(xh) if history else null, _p(wh) if history else null, _p(weh) if history else null)
        out.update(ll=ll, x=xh, w=wh, we=weh)
        return out

    def loglik(self, u, y, epoch=0):
        u, y = _f64(u).reshape(-1, max(self.nu, 1))[:, : self.nu], _f64(y).reshape(-1, self.ny)
        T = y.shape[0]
        uu = np.ascontiguousarray(u) if self.nu else np.zeros((T, 1))
        out = dict(ll_steps=np.zeros(T), ess=np.zeros(T), resampled=np.zeros(T, dtype=np.int32))
        ll = lib().orc_loglik(self.h, T, _p(uu), _p(y), epoch, _p(out["ll_steps"]), _p(out["ess"]),
                              out["resampled"].ctypes.data_as(i32p))
        out["ll"] = ll
        return out

    def set_user_functions(self, dynamics=None, loglik=None):
        """dynamics(x, u, t) -> x_next (no noise) and / or loglik(x, u, y, t) -> float as Python callables replacing the
        descriptor model (the reference takes arbitrary closures, PFtypes.jl:128,232).  Slow (one Python call per
        particle): for parity tests of user-supplied device functions at small N."""
        nx, nu, ny = self.nx, self.nu, self.ny
        DYN = C.CFUNCTYPE(None, dp, dp, C.c_double, dp, C.c_void_p)
        LIK = C.CFUNCTYPE(C.c_double, dp, dp, dp, C.c_double, C.c_void_p)

        def _arr(p, n):
            return np.ctypeslib.as_array(p, shape=(n,)) if n else np.zeros(0)

        def dyn_cb(xp, up, t, outp, _ctx):
            out = np.asarray(dynamics(_arr(xp, nx).copy(), _arr(up, nu).copy(), t), dtype=np.float64)
            _arr(outp, nx)[:] = out

        def lik_cb(xp, up, yp, t, _ctx):
            return float(loglik(_arr(xp, nx).copy(), _arr(up, nu).copy(), _arr(yp, ny).copy(), t))

        self._cb_dyn = DYN(dyn_cb) if dynamics is not None else None      # keep the thunks alive
        self._cb_lik = LIK(lik_cb) if loglik is not None else None
        rc = lib().orc_set_user_functions(self.h, C.cast(self._cb_dyn, C.c_void_p) if self._cb_dyn else None,
                                          C.cast(self._cb_lik, C.c_void_p) if self._cb_lik else None, None)
        if rc:
            raise RuntimeError(f"orc_set_user_functions failed: {rc}")

    def smooth(self, M, u, xf, wf, wef, epoch=0):
        """smooth(pf, xf, wf, wef, ll, M, u, y)  smoothing.jl:116-143 -> xb [T][M][nx]"""
        xf, wf, wef = _f64(xf), _f64(wf), _f64(wef)
        T = wf.reshape(-1, self.N).shape[0]
        u = _f64(u).reshape(-1, max(self.nu, 1))
        xb = np.zeros((T, M, self.nx))
        rc = lib().orc_smooth(self.h, T, M, _p(np.ascontiguousarray(u)), _p(xf), _p(wf), _p(wef), epoch, _p(xb))
        if rc:
            raise RuntimeError(f"orc_smooth failed: {rc}")
        return xb

    def simulate(self, u, sim_seed=1):
        u = _f64(u).reshape(-1, max(self.nu, 1))
        T = u.shape[0]
        xs = np.zeros((T, self.nx))
        ys = np.zeros((T, self.ny))
        lib().orc_simulate(self.h, T, _p(u), sim_seed, _p(xs), _p(ys))
        return xs, ys

    def dynamics(self, x, u, t):
        x, u = _f64(x), self._u(u)
        out = np.zeros(self.nx)
        lib().orc_dynamics(self.h, _p(x), _p(u), float(t), _p(out))
        return out


def kalman_loglik(model, u, y):
    u, y = _f64(u).reshape(-1, max(model.nu, 1)), _f64(y).reshape(-1, model.ny)
    return lib().orc_kalman_loglik(C.byref(model.struct), y.shape[0], _p(u), _p(y))
