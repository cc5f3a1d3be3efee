"""Ensemble Kalman filter on the device — host side (reference src/enkf.jl, stochastic EnKF with perturbed observations).

`EnsembleKalmanFilter(dynamics, measurement, R1, R2, d0, N; nu, ny, Ts, inflation)` mirrors enkf.jl:94-141 with descriptor
arguments (closures cannot cross the C-ABI): dynamics = LinearDynamics | QuadtankRK4, measurement = LinearMeasurement.
The ensemble is the particle buffer of an ordinary handle; the verbs are the llpf_enkf_* entry points of include/llpf.h
(kernel: csrc/llpf_enkf.cuh).  `seed` replaces `rng` (ensemble stream 0, process noise 1, observation
perturbations 8, all keyed by (seed, epoch; step = enkf.t, member))."""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi, filters as F
from ._abi import check

dp = C.POINTER(C.c_double)


@dataclass
class KalmanFilteringSolution:  # src/solutions.jl: KalmanFilteringSolution(f,u,y,x,xt,R,Rt,ll,e,K,S,...)
    f: object
    u: np.ndarray
    y: np.ndarray
    x: np.ndarray      # [T][nx]      predictions x(t|t-1)
    xt: np.ndarray     # [T][nx]      filtered    x(t|t)
    R: np.ndarray      # [T][nx][nx]
    Rt: np.ndarray     # [T][nx][nx]
    ll: float
    e: np.ndarray      # [T][ny]      innovations
    K: np.ndarray      # [T][nx][ny]
    S: np.ndarray      # [T][ny][ny]
    t: np.ndarray
    extra: dict = field(default_factory=dict)


class EnsembleKalmanFilter:
    def __init__(self, dynamics, measurement, R1, R2, d0, N, *, nu=None, ny=None, p=None, Ts=1.0, inflation=1.0, seed=0,
                 device=0, **_ignored):
        if not isinstance(measurement, F.LinearMeasurement):
            raise TypeError("measurement must be a LinearMeasurement descriptor")
        self._lib = _abi.load_library()
        R1, R2 = np.atleast_2d(np.asarray(R1, dtype=np.float64)), np.atleast_2d(np.asarray(R2, dtype=np.float64))
        self._model = F._ModelBuffers(dynamics, measurement.C, R1, R2, d0)
        self.dynamics, self.measurement, self.R1, self.R2, self.d0 = dynamics, measurement, R1, R2, d0
        self.nx, self.nu, self.ny = self._model.nx, self._model.nu, self._model.ny
        if nu is not None and int(nu) != self.nu:
            raise ValueError(f"nu = {nu} does not match the dynamics descriptor ({self.nu})")
        cfg = _abi.Config()
        cfg.N, cfg.filter, cfg.resampling = int(N), _abi.FILTER_PF, F.ResampleSystematic.code
        cfg.resample_threshold, cfg.Ts, cfg.seed = 0.0, float(Ts), int(seed)
        cfg.scan_mode, cfg.device, cfg.rank, cfg.world = _abi.SCAN_FAST, int(device), 0, 1
        cfg.particle_dtype = _abi.PARTICLE_F64
        self._cfg = cfg
        self._h = C.c_void_p()
        check(self._lib, self._lib.llpf_create(C.byref(cfg), C.byref(self._model.struct), C.byref(self._h)))
        self.N, self.Ts, self.seed, self.p = int(N), float(Ts), int(seed), p
        self.inflation = float(inflation)
        if self.inflation < 1.0:
            import warnings
            warnings.warn("Inflation factor should be ≥ 1.0 to prevent filter divergence.")   # enkf.jl:119
        check(self._lib, self._lib.llpf_enkf_set_inflation(self._h, self.inflation))
        self._epoch = 0
        check(self._lib, self._lib.llpf_enkf_reset(self._h, 0))      # the constructor draws the ensemble  enkf.jl:121-124

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.llpf_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- accessors  enkf.jl:178-193
    def _state(self):
        m, c, t = np.zeros(self.nx), np.zeros((self.nx, self.nx)), C.c_int64()
        check(self._lib, self._lib.llpf_enkf_state(self._h, m.ctypes.data_as(dp), c.ctypes.data_as(dp), C.byref(t)))
        return m, c, int(t.value)

    @property
    def t(self):
        return self._state()[2]

    def __call__(self, u, y, p=None, t=None):       # (enkf::EnsembleKalmanFilter)(u, y, p, t)  enkf.jl:369
        return enkf_update(self, u, y, p, t)


def _vec(v, n):
    if n == 0:
        return None, C.cast(None, dp)
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
    if a.size != n:
        raise ValueError(f"vector has length {a.size}, expected {n}")
    return a, a.ctypes.data_as(dp)


def enkf_state(enkf):
    """state(enkf): the cached ensemble mean  enkf.jl:186"""
    return enkf._state()[0]


def enkf_covariance(enkf):
    """covariance(enkf): the cached sample covariance  enkf.jl:193"""
    return enkf._state()[1]


def enkf_particles(enkf):
    """particles(enkf): the ensemble [N][nx]"""
    out = np.zeros((enkf.N, enkf.nx))
    check(enkf._lib, enkf._lib.llpf_get_particles(enkf._h, out.ctypes.data_as(dp)))
    return out


def enkf_reset(enkf, epoch=None):
    """reset!(enkf)  enkf.jl:205-224"""
    if epoch is None:
        enkf._epoch += 1
        epoch = enkf._epoch
    else:
        enkf._epoch = int(epoch)
    check(enkf._lib, enkf._lib.llpf_enkf_reset(enkf._h, int(epoch)))


def enkf_predict(enkf, u, p=None, t=None):
    """predict!(enkf, u, p, t)  enkf.jl:228-272"""
    _, up = _vec(u, enkf.nu)
    t = enkf.t * enkf.Ts if t is None else float(t)
    check(enkf._lib, enkf._lib.llpf_enkf_predict(enkf._h, up, t))


def enkf_correct(enkf, u, y, p=None, t=None):
    """correct!(enkf, u, y, p, t) -> dict(ll, e, S, K)  enkf.jl:281-356"""
    _, up = _vec(u, enkf.nu)
    _, yp = _vec(y, enkf.ny)
    t = enkf.t * enkf.Ts if t is None else float(t)
    ll = C.c_double()
    e, S, K = np.zeros(enkf.ny), np.zeros((enkf.ny, enkf.ny)), np.zeros((enkf.nx, enkf.ny))
    llv = np.zeros(1)
    check(enkf._lib, enkf._lib.llpf_enkf_correct(enkf._h, up, yp, t, llv.ctypes.data_as(dp), e.ctypes.data_as(dp),
                                                S.ctypes.data_as(dp), K.ctypes.data_as(dp)))
    return dict(ll=float(llv[0]), e=e, S=S, K=K)


def enkf_update(enkf, u, y, p=None, t=None):
    """update!(enkf, u, y, p, t) = correct! then predict!  enkf.jl:361-366"""
    t = enkf.t * enkf.Ts if t is None else float(t)
    r = enkf_correct(enkf, u, y, p, t)
    enkf_predict(enkf, u, p, t)
    return r


def enkf_forward_trajectory(enkf, u, y, p=None, *, epoch=None):
    """forward_trajectory(enkf, u, y) -> KalmanFilteringSolution  filtering.jl:282-325, one launch"""
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1, enkf.ny))
    T = y.shape[0]
    if enkf.nu > 0:
        u = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(T, enkf.nu))
        up = u.ctypes.data_as(dp)
    else:
        u, up = None, C.cast(None, dp)
    if epoch is None:
        enkf._epoch += 1
        epoch = enkf._epoch
    else:
        enkf._epoch = int(epoch)
    nx, ny = enkf.nx, enkf.ny
    x, xt = np.zeros((T, nx)), np.zeros((T, nx))
    R, Rt = np.zeros((T, nx, nx)), np.zeros((T, nx, nx))
    e, lls, S, K = np.zeros((T, ny)), np.zeros(T), np.zeros((T, ny, ny)), np.zeros((T, nx, ny))
    ll = C.c_double()
    P = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    check(enkf._lib, enkf._lib.llpf_enkf_run(enkf._h, T, up, P(y), int(epoch), C.cast(C.byref(ll), dp), P(x), P(R), P(xt),
                                            P(Rt), P(e), P(lls), P(S), P(K)))
    return KalmanFilteringSolution(enkf, u, y, x, xt, R, Rt, float(ll.value), e, K, S, np.arange(T) * enkf.Ts,
                                   dict(ll_steps=lls))
