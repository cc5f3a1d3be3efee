"""Host-side mirror of the reference's particle-filter API for the accelerated path.

Same names, argument order and return conventions as LowLevelParticleFilters.jl (Julia `f!` -> `f`):

    ParticleFilter(N, dynamics, measurement, dynamics_density, measurement_density, initial_density; kw...)
        src/PFtypes.jl:65-75        kw: resample_threshold=0.1, resampling_strategy=ResampleSystematic, Ts=1.0
    AdvancedParticleFilter(N, dynamics, measurement, measurement_likelihood, dynamics_density, initial_density; kw...)
        src/PFtypes.jl:200-210      kw: resample_threshold=0.5
    AuxiliaryParticleFilter(pf | args...)                         src/PFtypes.jl:38-49
    reset(pf) predict(pf,u,p,t) correct(pf,u,y,p,t) update(pf,u,y,p,t) pf(u,y)   src/filtering.jl:4-14,140-191,238-240
    forward_trajectory(pf,u,y) -> ParticleFilteringSolution       src/filtering.jl:343-384, src/solutions.jl:334-345
    loglik(pf,u,y)                                                src/smoothing.jl:227-236
    particles weights expweights num_particles effective_particles shouldresample weighted_mean index
                                                                  src/PFtypes.jl:296-334, src/resample.jl:1-10

User closures cannot cross the C-ABI, so `dynamics`, `measurement`, the densities and
`measurement_likelihood` are *descriptors* (LinearDynamics, QuadtankRK4, LinearMeasurement, MvNormal,
GaussianLikelihood).  `rng=` is replaced by `seed=` (counter-based Philox streams).
Everything runs through libllpf_b200.so; there is no CPU fallback.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from ._abi import (DYN_LINEAR, DYN_QUADTANK_RK4, DYN_USER, FILTER_ADVANCED, FILTER_AUX, FILTER_AUX_ADVANCED,
                   FILTER_PF, RESAMPLE_STRATIFIED, RESAMPLE_SYSTEMATIC, SCAN_FAST, SCAN_SERIAL,
                   TIME_FORWARD_TRAJECTORY, TIME_LOGLIK, check)

dp = _abi.c_double_p


# ---------------------------------------------------------------------------------------------
# descriptors
# ---------------------------------------------------------------------------------------------
class ResamplingStrategy:  # src/LowLevelParticleFilters.jl:43-46
    code = None


class ResampleSystematic(ResamplingStrategy):
    code = RESAMPLE_SYSTEMATIC


class ResampleStratified(ResamplingStrategy):
    code = RESAMPLE_STRATIFIED


class ResampleResidual(ResamplingStrategy):
    code = _abi.RESAMPLE_RESIDUAL


class ResampleMetropolis(ResamplingStrategy):
    """NOT in the reference: Metropolis resampling (Murray, Lee & Jacob 2016) — per output slot a B-step Metropolis chain
    over the particle indices using weight ratios only (no prefix sum, one grid barrier).  Biased for finite B
    (`metropolis_steps=` of the filter constructors, default 32).  Single-GPU, Float64 particles."""
    code = _abi.RESAMPLE_METROPOLIS


@dataclass
class MvNormal:
    """MvNormal(mu, Sigma) / MvNormal(Sigma) — Distributions.MvNormal or SimpleMvNormal (src/utils.jl:241-273)."""
    mu: np.ndarray
    Sigma: np.ndarray = None
    dtype: object = np.float64   # element type of the samples: an `initial_density` of Float32 makes Float32 particles (PFtypes.jl:66)

    def __post_init__(self):
        self.dtype = np.dtype(self.dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise TypeError("MvNormal dtype must be float64 or float32")
        if self.Sigma is None:
            self.Sigma = np.atleast_2d(np.asarray(self.mu, dtype=np.float64))
            self.mu = np.zeros(self.Sigma.shape[0])
        self.mu = np.asarray(self.mu, dtype=np.float64).reshape(-1)
        S = np.asarray(self.Sigma, dtype=np.float64)
        if S.ndim == 0:
            S = float(S) * np.eye(self.mu.size)
        elif S.ndim == 1:
            S = np.diag(S)
        self.Sigma = S

    def __len__(self):
        return self.mu.size


@dataclass
class LinearDynamics:
    """dynamics(x,u,p,t) = A*x .+ B*u   (examples/example_lineargaussian.jl:27)"""
    A: np.ndarray
    B: np.ndarray = None


@dataclass
class QuadtankRK4:
    """rk4(quadtank, Ts; supersample) — examples/example_quadtank.jl:91-106,35 ; src/utils.jl:220-237.
    p = [kc, k1, k2, A, a, gamma]; t_switch/a1_factor reproduce `if t > 500; a1 *= 2` (:15-17)."""
    p: tuple = (0.5, 1.6, 1.6, 4.9, 0.03, 0.2)
    Ts: float = 1.0
    supersample: int = 2
    t_switch: float = float("inf")
    a1_factor: float = 1.0


@dataclass
class LinearMeasurement:
    """measurement(x,u,p,t) = C*x"""
    C: np.ndarray


@dataclass
class GaussianLikelihood:
    """measurement_likelihood(x,u,y,p,t) = logpdf(MvNormal(R2), C*x - y)  (example_lineargaussian.jl:238-240)"""
    C: np.ndarray
    R2: np.ndarray


# ---- user-defined models: the reference's closures, given as CUDA device code (llpf_create_user) ----------------
@dataclass
class CudaDynamics:
    """dynamics(x,u,p,t) (PFtypes.jl:128,255) as the BODY of
        __device__ void dynamics(double (&x)[NX], const double* u, const double* p, double t)
    It must overwrite x with f(x,u,p,t) WITHOUT noise: the additive N(0,R1) of `dynamics_density` is drawn by the engine
    (PFtypes.jl:135).  `nu` = length of u.  Compiled at run time (NVRTC) and inlined into the fused sweep."""
    body: str
    nu: int = 0


@dataclass
class CudaLikelihood:
    """measurement_likelihood(x,u,y,p,t) (PFtypes.jl:232) as the BODY of
        __device__ double loglik(const double (&x)[NX], const double* u, const double* y, const double* p, double t)
    returning the log-likelihood of y given x.  `ny` = length of y."""
    body: str
    ny: int = 1


@dataclass
class CudaMeasurement:
    """measurement(x,u,p,t) (PFtypes.jl:116) as the BODY of a function that writes the predicted measurement into
    `double yh[NY]` (x, u, p, t in scope); the Gaussian `measurement_density` supplies logpdf(dg, y - yh)."""
    body: str


def _lit(v):
    return float(v).hex()


def _linear_dynamics_body(A, B, nx, nu):
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    lines = [f"double xn[{nx}];"]
    for r in range(nx):
        terms = " + ".join(f"{_lit(A[r, c])} * x[{c}]" for c in range(nx))
        if nu:
            Bm = np.asarray(B, dtype=np.float64).reshape(nx, nu)
            terms = f"({terms}) + (" + " + ".join(f"{_lit(Bm[r, c])} * u[{c}]" for c in range(nu)) + ")"
        lines.append(f"xn[{r}] = {terms};")
    lines += [f"x[{r}] = xn[{r}];" for r in range(nx)]
    return "\n".join(lines)


def _gaussian_loglik_body(meas_body, R2, ny):
    """logpdf(N(0,R2), y - yh) = c0 - |L \\ (y - yh)|^2 / 2  (utils.jl:252-257), L = chol(R2) as literals"""
    R2 = np.atleast_2d(np.asarray(R2, dtype=np.float64))
    L = np.linalg.cholesky(R2)
    c0 = -(ny * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(L)))) / 2
    lines = [f"double yh[{ny}];", "{", meas_body, "}", f"double v[{ny}]; double q = 0.0;"]
    for i in range(ny):
        acc = f"(y[{i}] - yh[{i}])" + "".join(f" - {_lit(L[i, k])} * v[{k}]" for k in range(i))
        lines.append(f"v[{i}] = ({acc}) / {_lit(L[i, i])}; q += v[{i}] * v[{i}];")
    lines.append(f"return {_lit(c0)} - q / 2;")
    return "\n".join(lines)


def _user_source(nx, dyn_body, lik_body):
    return (f"namespace llpf_user {{\n"
            f"template <> __device__ void dynamics<{nx}>(double (&x)[{nx}], const double* u, const double* p, double t) {{\n"
            f"{dyn_body}\n}}\n"
            f"template <> __device__ double loglik<{nx}>(const double (&x)[{nx}], const double* u, const double* y, "
            f"const double* p, double t) {{\n{lik_body}\n}}\n}}  // namespace llpf_user\n")


class _UserModelBuffers:
    """Model struct + generated CUDA source of a filter whose dynamics and / or likelihood are device code."""

    def __init__(self, dynamics, likelihood_body, ny, R1, d0, source=None, nu=None):
        nx = len(d0)
        if source is not None:          # complete llpf_user source generated elsewhere (rbpf.py)
            nu, dyn_body = int(nu or 0), None
        elif isinstance(dynamics, CudaDynamics):
            nu, dyn_body = int(dynamics.nu), dynamics.body
        elif isinstance(dynamics, LinearDynamics):
            nu = 0 if dynamics.B is None else np.asarray(dynamics.B).reshape(nx, -1).shape[1]
            dyn_body = _linear_dynamics_body(dynamics.A, dynamics.B, nx, nu)
        else:
            raise TypeError("with user-defined device code the dynamics must be CudaDynamics or LinearDynamics")
        self.source = source if source is not None else _user_source(nx, dyn_body, likelihood_body)
        m = _abi.Model()
        self.keep = []
        null = C.cast(None, dp)

        def put(M, shape):
            a = _colmajor(M, shape)
            self.keep.append(a)
            return a.ctypes.data_as(dp)

        m.nx, m.nu, m.ny, m.dynamics = nx, nu, int(ny), DYN_USER
        m.A = m.B = m.C = m.R2 = null
        m.R1 = put(R1, (nx, nx))
        m.mu0 = put(d0.mu, (nx,))
        m.Sigma0 = put(d0.Sigma, (nx, nx))
        m.t_switch, m.a1_factor, m.integ_Ts, m.supersample = float("inf"), 1.0, 1.0, 1
        self.struct = m
        self.nx, self.nu, self.ny = nx, nu, int(ny)


@dataclass
class ParticleFilteringSolution:  # src/solutions.jl:334-345
    f: object
    u: np.ndarray
    y: np.ndarray
    x: np.ndarray      # [T][N][nx]  (reference: N x T Matrix{SVector}; same memory order)
    w: np.ndarray      # [T][N]
    we: np.ndarray     # [T][N]
    ll: float
    t: np.ndarray
    extra: dict = field(default_factory=dict)


def _colmajor(M, shape):
    return np.asfortranarray(np.asarray(M, dtype=np.float64).reshape(shape))


class _ModelBuffers:
    """Keeps the column-major arrays alive for as long as the ctypes struct points at them."""

    def __init__(self, dynamics, C_, R1, R2, d0):
        C_ = np.atleast_2d(np.asarray(C_, dtype=np.float64))
        ny, nx = C_.shape
        m = _abi.Model()
        self.keep = []
        null = C.cast(None, dp)

        def put(M, shape):
            a = _colmajor(M, shape)
            self.keep.append(a)
            return a.ctypes.data_as(dp)

        if isinstance(dynamics, LinearDynamics):
            A = np.atleast_2d(np.asarray(dynamics.A, dtype=np.float64))
            nu = 0 if dynamics.B is None else np.atleast_2d(np.asarray(dynamics.B)).reshape(nx, -1).shape[1]
            m.dynamics = DYN_LINEAR
            m.A = put(A, (nx, nx))
            m.B = put(dynamics.B, (nx, nu)) if nu > 0 else null
        elif isinstance(dynamics, QuadtankRK4):
            nu = 2
            m.dynamics = DYN_QUADTANK_RK4
            m.A, m.B = null, null
            for k, v in enumerate(dynamics.p):
                m.dyn_params[k] = float(v)
            m.t_switch, m.a1_factor = float(dynamics.t_switch), float(dynamics.a1_factor)
            m.integ_Ts, m.supersample = float(dynamics.Ts), int(dynamics.supersample)
        else:
            raise TypeError("dynamics must be a descriptor (LinearDynamics | QuadtankRK4); closures cannot cross the C-ABI")
        m.nx, m.nu, m.ny = nx, nu, ny
        m.C = put(C_, (ny, nx))
        m.R1 = put(R1, (nx, nx))
        m.R2 = put(R2, (ny, ny))
        m.mu0 = put(d0.mu, (nx,))
        m.Sigma0 = put(d0.Sigma, (nx, nx))
        self.struct = m
        self.nx, self.nu, self.ny = nx, nu, ny


# ---------------------------------------------------------------------------------------------
# filters
# ---------------------------------------------------------------------------------------------
class AbstractParticleFilter:
    _filter_code = FILTER_PF

    def _create(self, N, model, resample_threshold, resampling_strategy, Ts, seed, scan_mode, device,
                rank=0, world=1, particle_dtype=np.float64, p=None, single_block=False, metropolis_steps=0):
        """N is the GLOBAL particle count; with world > 1 this process owns the contiguous slice
        [rank*N/world, (rank+1)*N/world) and must call connect_shards() before the first step."""
        self._lib = _abi.load_library()
        self._model = model
        cfg = _abi.Config()
        cfg.N = int(N)
        cfg.filter = self._filter_code
        cfg.resampling = resampling_strategy.code
        cfg.resample_threshold = float(resample_threshold)
        cfg.Ts = float(Ts)
        cfg.seed = int(seed)
        cfg.scan_mode = SCAN_SERIAL if scan_mode in (SCAN_SERIAL, "serial") else SCAN_FAST
        cfg.device, cfg.rank, cfg.world = int(device), int(rank), int(world)
        self.particle_dtype = np.dtype(particle_dtype)
        cfg.particle_dtype = _abi.PARTICLE_F32 if self.particle_dtype == np.dtype(np.float32) else _abi.PARTICLE_F64
        cfg.single_block = 1 if single_block else 0
        cfg.metropolis_steps = int(metropolis_steps)
        self.single_block = bool(single_block)
        self._cfg = cfg
        self._h = C.c_void_p()
        self.p = None
        if isinstance(model, _UserModelBuffers):
            pv = np.ascontiguousarray(np.asarray([] if p is None else p, dtype=np.float64).reshape(-1))
            self.p = pv
            check(self._lib, self._lib.llpf_create_user(C.byref(cfg), C.byref(model.struct), model.source.encode(),
                                                        pv.ctypes.data_as(dp) if pv.size else C.cast(None, dp),
                                                        int(pv.size), C.byref(self._h)))
        else:
            check(self._lib, self._lib.llpf_create(C.byref(cfg), C.byref(model.struct), C.byref(self._h)))
        self.N_global = int(N)
        self.rank, self.world = int(rank), max(1, int(world))
        self.N = int(N) // self.world          # local particle count: accessor arrays have this length
        self.first = self.rank * self.N
        self.nx, self.nu, self.ny = model.nx, model.nu, model.ny
        self.Ts = float(Ts)
        self.resample_threshold = float(resample_threshold)
        self.resampling_strategy = resampling_strategy
        self.seed = int(seed)
        self._epoch = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.llpf_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- helpers
    def _vec(self, v, n, name):
        if n == 0:
            return None, C.cast(None, dp)
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
        if a.size != n:
            raise ValueError(f"{name} has length {a.size}, expected {n}")
        return a, a.ctypes.data_as(dp)

    def _default_t(self, t):
        return index(self) * self.Ts if t is None else float(t)

    def _use_p(self, p):
        """Per-call parameter override `p` (filtering.jl:140,164).  User-defined models: the vector handed to the device
        functions is replaced (no recompilation).  Descriptor models carry their parameters in the matrices: use
        set_model(pf, ...) — silently ignoring `p` would compute with the wrong parameters, so it raises."""
        if p is None:
            return
        if isinstance(self._model, _UserModelBuffers):
            pv = np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(-1))
            check(self._lib, self._lib.llpf_set_user_params(self._h, pv.ctypes.data_as(dp) if pv.size else C.cast(None, dp),
                                                            int(pv.size)))
            self.p = pv
            return
        raise TypeError("descriptor models have no parameter vector: change the model with set_model(pf, ...) "
                        "(the reference's `p` override, filtering.jl:140,164, applies to closures)")

    def __call__(self, u, y, p=None, t=None):  # (pf::ParticleFilter)(u,y,p,t) filtering.jl:238
        return update(self, u, y, p, t)


class ParticleFilter(AbstractParticleFilter):
    _filter_code = FILTER_PF

    def __init__(self, N, dynamics, measurement, dynamics_density, measurement_density, initial_density, *,
                 resample_threshold=0.1, resampling_strategy=ResampleSystematic, Ts=1.0, seed=0,
                 scan_mode="fast", device=0, p=None, rank=0, world=1, single_block=False, metropolis_steps=0,
                 **_ignored):
        self.dynamics, self.measurement = dynamics, measurement
        self.dynamics_density, self.measurement_density = dynamics_density, measurement_density
        self.initial_density = initial_density
        ny = len(measurement_density)
        if isinstance(dynamics, CudaDynamics) or isinstance(measurement, CudaMeasurement):
            # closures of the reference (PFtypes.jl:116,128) given as device code: w[i] += logpdf(dg, y - g(x[i]))
            if isinstance(measurement, CudaMeasurement):
                meas_body = measurement.body
            elif isinstance(measurement, LinearMeasurement):
                Cm = np.atleast_2d(np.asarray(measurement.C, dtype=np.float64))
                meas_body = "\n".join(f"yh[{a}] = " + " + ".join(f"{_lit(Cm[a, c])} * x[{c}]" for c in range(Cm.shape[1])) + ";"
                                      for a in range(ny))
            else:
                raise TypeError("measurement must be a LinearMeasurement or CudaMeasurement descriptor")
            lik = _gaussian_loglik_body(meas_body, measurement_density.Sigma, ny)
            model = _UserModelBuffers(dynamics, lik, ny, dynamics_density.Sigma, initial_density)
        else:
            if not isinstance(measurement, LinearMeasurement):
                raise TypeError("measurement must be a LinearMeasurement or CudaMeasurement descriptor")
            model = _ModelBuffers(dynamics, measurement.C, dynamics_density.Sigma, measurement_density.Sigma,
                                  initial_density)
        self._create(N, model, resample_threshold, resampling_strategy, Ts, seed, scan_mode, device, rank, world,
                     particle_dtype=getattr(initial_density, "dtype", np.float64), p=p, single_block=single_block, metropolis_steps=metropolis_steps)


class AdvancedParticleFilter(AbstractParticleFilter):
    _filter_code = FILTER_ADVANCED

    def __init__(self, N, dynamics, measurement, measurement_likelihood, dynamics_density, initial_density, *,
                 resample_threshold=0.5, resampling_strategy=ResampleSystematic, Ts=1.0, seed=0,
                 scan_mode="fast", device=0, p=None, rank=0, world=1, single_block=False, metropolis_steps=0,
                 **_ignored):
        self.dynamics, self.measurement = dynamics, measurement
        self.measurement_likelihood = measurement_likelihood
        self.dynamics_density, self.initial_density = dynamics_density, initial_density
        if isinstance(dynamics, CudaDynamics) or isinstance(measurement_likelihood, CudaLikelihood):
            # the reference's closures dynamics(x,u,p,t,noise) / measurement_likelihood(x,u,y,p,t) (PFtypes.jl:232,255) as
            # device code; the noise stays additive Gaussian (dynamics_density), drawn by the engine
            if isinstance(measurement_likelihood, CudaLikelihood):
                lik, ny = measurement_likelihood.body, int(measurement_likelihood.ny)
            elif isinstance(measurement_likelihood, GaussianLikelihood):
                Cm = np.atleast_2d(np.asarray(measurement_likelihood.C, dtype=np.float64))
                ny = Cm.shape[0]
                meas_body = "\n".join(f"yh[{a}] = " + " + ".join(f"{_lit(Cm[a, c])} * x[{c}]" for c in range(Cm.shape[1])) + ";"
                                      for a in range(ny))
                lik = _gaussian_loglik_body(meas_body, measurement_likelihood.R2, ny)
            else:
                raise TypeError("measurement_likelihood must be a GaussianLikelihood or CudaLikelihood descriptor")
            model = _UserModelBuffers(dynamics, lik, ny, dynamics_density.Sigma, initial_density)
        else:
            if not isinstance(measurement_likelihood, GaussianLikelihood):
                raise TypeError("measurement_likelihood must be a GaussianLikelihood or CudaLikelihood descriptor")
            model = _ModelBuffers(dynamics, measurement_likelihood.C, dynamics_density.Sigma,
                                  measurement_likelihood.R2, initial_density)
        self._create(N, model, resample_threshold, resampling_strategy, Ts, seed, scan_mode, device, rank, world,
                     particle_dtype=getattr(initial_density, "dtype", np.float64), p=p, single_block=single_block, metropolis_steps=metropolis_steps)


class AuxiliaryParticleFilter(AbstractParticleFilter):
    """AuxiliaryParticleFilter(pf) or AuxiliaryParticleFilter(args...; kwargs...)  PFtypes.jl:38-49"""

    def __init__(self, *args, **kwargs):
        if len(args) == 1 and isinstance(args[0], AbstractParticleFilter):
            inner = args[0]
            adv = isinstance(inner, AdvancedParticleFilter)
            cfg = inner._cfg
            self.pf = inner
            self._filter_code = FILTER_AUX_ADVANCED if adv else FILTER_AUX
            for name in ("dynamics", "measurement", "dynamics_density", "initial_density"):
                setattr(self, name, getattr(inner, name))
            self._create(inner.N_global, inner._model, inner.resample_threshold, inner.resampling_strategy, inner.Ts,
                         inner.seed, cfg.scan_mode, cfg.device, inner.rank, inner.world,
                         particle_dtype=inner.particle_dtype, p=inner.p, single_block=inner.single_block,
                         metropolis_steps=inner._cfg.metropolis_steps)
        else:
            self.__init__(ParticleFilter(*args, **kwargs))

    def __call__(self, u, y, y1, p=None, t=None):  # (pf::AuxiliaryParticleFilter)(u,y,y1,p,t) filtering.jl:239
        return update(self, u, y, p, t, y1=y1)


# ---------------------------------------------------------------------------------------------
# verbs
# ---------------------------------------------------------------------------------------------
def reset(pf, epoch=None):
    """reset!(pf)  filtering.jl:4-14.  Successive calls advance to a fresh RNG epoch unless one is given."""
    if epoch is None:
        pf._epoch += 1
        epoch = pf._epoch
    else:
        pf._epoch = int(epoch)
    check(pf._lib, pf._lib.llpf_reset(pf._h, int(epoch)))


def correct(pf, u, y, p=None, t=None):
    """correct!(pf,u,y,p,t) -> (ll, 0)  filtering.jl:164-174"""
    pf._use_p(p)
    _, up = pf._vec(u, pf.nu, "u")
    ya, yp = pf._vec(y, pf.ny, "y")
    ll = C.c_double()
    check(pf._lib, pf._lib.llpf_correct(pf._h, up, yp, pf._default_t(t), C.byref(ll)))
    return ll.value, 0


def predict(pf, u, p=None, t=None, y1=None):
    """predict!(pf,u,p,t) filtering.jl:140-153 ; predict!(pfa,u,y1,p,t) :195-234"""
    pf._use_p(p)
    _, up = pf._vec(u, pf.nu, "u")
    if isinstance(pf, AuxiliaryParticleFilter):
        if y1 is None:
            raise TypeError("predict!(pfa, u, y1, p, t) needs y1")
        _, y1p = pf._vec(y1, pf.ny, "y1")
        check(pf._lib, pf._lib.llpf_predict_aux(pf._h, up, y1p, pf._default_t(t)))
    else:
        check(pf._lib, pf._lib.llpf_predict(pf._h, up, pf._default_t(t)))


def update(pf, u, y, p=None, t=None, y1=None):
    """update!(f,u,y,p,t) filtering.jl:181-185 ; update!(pfa,u,y,y1,p,t) :187-191.  Returns (ll, 0)."""
    pf._use_p(p)
    _, up = pf._vec(u, pf.nu, "u")
    _, yp = pf._vec(y, pf.ny, "y")
    null = C.cast(None, dp)
    y1p = null
    if isinstance(pf, AuxiliaryParticleFilter):
        if y1 is None:
            raise TypeError("update!(pfa, u, y, y1, p, t) needs y1")
        _, y1p = pf._vec(y1, pf.ny, "y1")
    ll = C.c_double()
    check(pf._lib, pf._lib.llpf_update(pf._h, up, yp, y1p, pf._default_t(t), C.byref(ll)))
    return ll.value, 0


def _traj_inputs(pf, u, y):
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1, pf.ny))
    T = y.shape[0]
    if pf.nu > 0:
        u = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(-1, pf.nu))
        if u.shape[0] != T:
            raise ValueError("u and y must have the same number of time steps")
        up = u.ctypes.data_as(dp)
    else:
        u, up = None, C.cast(None, dp)
    return u, up, y, y.ctypes.data_as(dp), T


def _run(pf, u, y, conv, history, epoch, want_steps=True, want_xhat=True):
    u, up, y, yp, T = _traj_inputs(pf, u, y)
    if epoch is None:
        pf._epoch += 1
        epoch = pf._epoch
    else:
        pf._epoch = int(epoch)
    out = _abi.RunOutputs()
    res = {}
    null = C.cast(None, dp)
    if want_steps:
        res["ll_steps"] = np.zeros(T)
        res["ess"] = np.zeros(T)
        res["resampled"] = np.zeros(T, dtype=np.int32)
        out.ll_steps = res["ll_steps"].ctypes.data_as(dp)
        out.ess_steps = res["ess"].ctypes.data_as(dp)
        out.resampled = res["resampled"].ctypes.data_as(_abi.c_int32_p)
    if want_xhat:
        res["xhat"] = np.zeros((T, pf.nx))
        out.xhat = res["xhat"].ctypes.data_as(dp)
    if history:
        # the device history buffers are indexed by GLOBAL particle number; a sharded rank fills its slice
        res["x"] = np.zeros((T, pf.N_global, pf.nx))
        res["w"] = np.zeros((T, pf.N_global))
        res["we"] = np.zeros((T, pf.N_global))
        out.x_hist = res["x"].ctypes.data_as(dp)
        out.w_hist = res["w"].ctypes.data_as(dp)
        out.we_hist = res["we"].ctypes.data_as(dp)
    else:
        out.x_hist = out.w_hist = out.we_hist = null
    ll = C.c_double()
    check(pf._lib, pf._lib.llpf_run(pf._h, T, up, yp, conv, int(epoch), C.byref(ll), C.byref(out)))
    res["ll"] = ll.value
    res["u"], res["y"], res["T"] = u, y, T
    if history and pf.world > 1:
        sl = slice(pf.first, pf.first + pf.N)
        res["x"], res["w"], res["we"] = res["x"][:, sl], res["w"][:, sl], res["we"][:, sl]
    return res


def forward_trajectory(pf, u, y, p=None, *, history=True, epoch=None, pre_correct_cb=None, post_correct_cb=None,
                       pre_predict_cb=None, post_predict_cb=None):
    """forward_trajectory(pf,u,y,p; pre_correct_cb, post_correct_cb, pre_predict_cb, post_predict_cb)
    -> ParticleFilteringSolution   filtering.jl:343-384.
    history=False skips the N x T x/w/we arrays (they are then None); the per-step ll, ESS, resample flags and weighted
    means are always returned in `sol.extra`.
    Without callbacks the whole trajectory is ONE kernel launch.  With any of the four callbacks of filtering.jl:343 the
    loop runs step by step through the verbs (one launch per correct! / predict!), calling them exactly where the
    reference does (:353-362): pre_correct_cb(pf,u,y,p,t), post_correct_cb(pf,u,y,p,t,ll), pre_predict_cb(pf,u,y,p,t,ll),
    post_predict_cb(pf,u,y,p,t).  The callbacks may inspect or modify the filter through the accessors / set_state.
    The result is the same trajectory (same RNG counters) as the fused launch."""
    pf._use_p(p)
    cbs = (pre_correct_cb, post_correct_cb, pre_predict_cb, post_predict_cb)
    if any(cb is not None for cb in cbs):
        if isinstance(pf, AuxiliaryParticleFilter):
            raise TypeError("forward_trajectory(::AuxiliaryParticleFilter) takes no callbacks (filtering.jl:367)")
        return _forward_trajectory_stepwise(pf, u, y, p, history, epoch, *cbs)
    wide = pf.particle_dtype == np.dtype(np.float32)
    if wide and history:
        # the Float32 wide engine records no history inside its fused loop (llpf_run: LLPF_ERR_UNSUPPORTED): the solution
        # is assembled by the reference's own loop on the step verbs and accessors — same RNG counters, same trajectory
        sol = _forward_trajectory_stepwise(pf, u, y, p, True, epoch, None, None, None, None)
        sol.extra["xhat"] = np.einsum("tnd,tn->td", sol.x, sol.we)       # weighted_mean per step  filtering.jl:541-548
        return sol
    r = _run(pf, u, y, TIME_FORWARD_TRAJECTORY, history, epoch, want_xhat=not wide)
    t = np.arange(r["T"]) * pf.Ts  # range(0, step=Ts, length=T)  solutions.jl:345
    extra = {k: r.get(k) for k in ("ll_steps", "ess", "resampled", "xhat")}
    return ParticleFilteringSolution(pf, r["u"], r["y"], r.get("x"), r.get("w"), r.get("we"), r["ll"], t, extra)


def _forward_trajectory_stepwise(pf, u, y, p, history, epoch, pre_correct_cb, post_correct_cb, pre_predict_cb,
                                 post_predict_cb):
    """the reference's loop, filtering.jl:351-363, on the step verbs"""
    u_, _, y_, _, T = _traj_inputs(pf, u, y)
    nothing = lambda *a: None  # noqa: E731
    pre_correct_cb, post_correct_cb = pre_correct_cb or nothing, post_correct_cb or nothing
    pre_predict_cb, post_predict_cb = pre_predict_cb or nothing, post_predict_cb or nothing
    reset(pf, epoch)
    N = pf.N
    x = np.zeros((T, N, pf.nx)) if history else None
    w = np.zeros((T, N)) if history else None
    we = np.zeros((T, N)) if history else None
    ll = 0.0
    lls, res = np.zeros(T), np.zeros(T, dtype=np.int32)
    for t in range(T):
        ti = t * pf.Ts
        ut = u_[t] if u_ is not None else None
        pre_correct_cb(pf, ut, y_[t], p, ti)
        lli = correct(pf, ut, y_[t], None, ti)[0]
        post_correct_cb(pf, ut, y_[t], p, ti, lli)
        ll += lli
        lls[t] = lli
        if history:
            x[t], w[t], we[t] = particles(pf), weights(pf), expweights(pf)
        pre_predict_cb(pf, ut, y_[t], p, ti, lli)
        res[t] = 1 if shouldresample(pf) else 0
        predict(pf, ut, None, ti)
        post_predict_cb(pf, ut, y_[t], p, ti)
    tt = np.arange(T) * pf.Ts
    return ParticleFilteringSolution(pf, u_, y_, x, w, we, ll, tt, dict(ll_steps=lls, resampled=res, ess=None, xhat=None))


def loglik(pf, u, y, p=None, *, epoch=None, details=False):
    """loglik(pf,u,y,p)  smoothing.jl:227-236"""
    pf._use_p(p)
    r = _run(pf, u, y, TIME_LOGLIK, False, epoch, want_steps=details, want_xhat=False)
    return r if details else r["ll"]


def trajectory_statistics(pf, u, y, p=None, *, q=(), epoch=None, mean=True, mode=True, cov=True):
    """forward_trajectory(pf, u, y) reduced ON THE DEVICE to the statistics the reference computes from the solution on the
    host: mean_trajectory / mode_trajectory (filtering.jl:417-440), weighted_cov (:575-583), weighted_quantile (:592-595).
    The N x T history stays in HBM (llpf_run_stats); nothing of size N crosses PCIe.
    Returns dict(ll, mean [T][nx], mode [T][nx], cov [T][nx][nx], quantile [T][len(q)][nx], ll_steps, ess, resampled)."""
    pf._use_p(p)
    u, up, y, yp, T = _traj_inputs(pf, u, y)
    if epoch is None:
        pf._epoch += 1
        epoch = pf._epoch
    else:
        pf._epoch = int(epoch)
    out = _abi.RunOutputs()
    res = dict(ll_steps=np.zeros(T), ess=np.zeros(T), resampled=np.zeros(T, dtype=np.int32))
    out.ll_steps = res["ll_steps"].ctypes.data_as(dp)
    out.ess_steps = res["ess"].ctypes.data_as(dp)
    out.resampled = res["resampled"].ctypes.data_as(_abi.c_int32_p)
    st = _abi.HistStats()
    qv = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(-1))
    if mean:
        res["mean"] = np.zeros((T, pf.nx)); st.xmean = res["mean"].ctypes.data_as(dp)
    if mode:
        res["mode"] = np.zeros((T, pf.nx)); st.xmode = res["mode"].ctypes.data_as(dp)
    if cov:
        res["cov"] = np.zeros((T, pf.nx, pf.nx)); st.xcov = res["cov"].ctypes.data_as(dp)
    if qv.size:
        res["quantile"] = np.zeros((T, qv.size, pf.nx))
        st.q, st.nq, st.xquantile = qv.ctypes.data_as(dp), int(qv.size), res["quantile"].ctypes.data_as(dp)
    ll = C.c_double()
    check(pf._lib, pf._lib.llpf_run_stats(pf._h, T, up, yp, int(epoch), C.byref(ll), C.byref(out), C.byref(st)))
    res["ll"] = ll.value
    return res


def loglik_batch(pfs, u, y, epochs=None, conv=TIME_LOGLIK):
    """[loglik(pf, u, y) for pf in pfs] in ONE kernel launch (llpf_run_batch): one thread block per filter.  The filters —
    each its own model / seed, e.g. one per Markov chain of `metropolis_threaded` (smoothing.jl:335-347) — must have been
    created with single_block=True and share dimensions.  epochs: RNG epoch per filter (default: each filter's next one).
    Every value is bit-identical to loglik(pf, u, y, epoch=...) on the same filter."""
    pfs = list(pfs)
    pf0 = pfs[0]
    u, up, y, yp, T = _traj_inputs(pf0, u, y)
    if epochs is None:
        epochs = []
        for pf in pfs:
            pf._epoch += 1
            epochs.append(pf._epoch)
    else:
        epochs = [int(e) for e in epochs]
        for pf, e in zip(pfs, epochs):
            pf._epoch = e
    Cn = len(pfs)
    hs = (C.c_void_p * Cn)(*[pf._h for pf in pfs])
    ep = (C.c_uint64 * Cn)(*epochs)
    out = np.zeros(Cn)
    check(pf0._lib, pf0._lib.llpf_run_batch(Cn, hs, T, up, yp, conv, ep, out.ctypes.data_as(dp)))
    return out


def mean_trajectory(*args, **kw):
    """mean_trajectory(sol) / mean_trajectory(x, we)  filtering.jl:417,436-438 -> T x nx ;
    mean_trajectory(pf,u,y) filtering.jl:393 -> (xhat, ll) through reduce_trajectory (:421-434)."""
    if len(args) == 1 and isinstance(args[0], ParticleFilteringSolution):
        sol = args[0]
        if sol.x is None:
            return sol.extra["xhat"]
        return np.einsum("tnd,tn->td", sol.x, sol.we)
    if len(args) == 2:
        x, we = args
        return np.einsum("tnd,tn->td", np.asarray(x), np.asarray(we))
    pf, u, y = args[:3]
    # reduce_trajectory: correct!(u[1],y[1],t=0); then pf(u[t-1], y[t], t=(t-1)Ts) for t=2..T
    yv = np.asarray(y, dtype=np.float64).reshape(-1, pf.ny)
    uv = np.asarray(u, dtype=np.float64).reshape(yv.shape[0], -1)
    reset(pf, kw.get("epoch"))
    ll = correct(pf, uv[0], yv[0], None, 0.0)[0]
    xh = [weighted_mean(pf)]
    for t in range(1, yv.shape[0]):
        ll += update(pf, uv[t - 1], yv[t], None, t * pf.Ts)[0]
        xh.append(weighted_mean(pf))
    return np.array(xh), ll


def mode_trajectory(sol):
    """mode_trajectory(sol)  filtering.jl:427,434 — particle with the largest weight at each step"""
    idx = np.argmax(sol.we, axis=1)
    return sol.x[np.arange(sol.x.shape[0]), idx]


# ---------------------------------------------------------------------------------------------
# particle smoother (FFBS)  smoothing.jl:104-143, helpers :350-385
# ---------------------------------------------------------------------------------------------
def smooth(pf, *args, p=None, epoch=None):
    """xb, ll = smooth(pf, M, u, y)                       smoothing.jl:104-107
       xb, ll = smooth(pf, xf, wf, wef, ll, M, u, y)      smoothing.jl:116-143
    xb is [T][M][nx] (the reference's M x T matrix of state vectors, column-major).  The forward history never
    leaves the device in the first form."""
    if len(args) == 3:
        M, u, y = args
        u, up, y, yp, T = _traj_inputs(pf, u, y)
        if epoch is None:
            pf._epoch += 1
            epoch = pf._epoch
        else:
            pf._epoch = int(epoch)
        xb = np.zeros((T, int(M), pf.nx))
        ll = C.c_double()
        check(pf._lib, pf._lib.llpf_smooth(pf._h, T, up, yp, int(M), int(epoch), C.byref(ll), xb.ctypes.data_as(dp),
                                           None))
        return xb, ll.value
    if len(args) == 7:
        xf, wf, wef, ll, M, u, y = args
        xf = np.ascontiguousarray(np.asarray(xf, dtype=np.float64))
        wf = np.ascontiguousarray(np.asarray(wf, dtype=np.float64))
        wef = np.ascontiguousarray(np.asarray(wef, dtype=np.float64))
        T = wf.shape[0]
        if xf.shape != (T, pf.N_global, pf.nx) or wf.shape != (T, pf.N_global) or wef.shape != wf.shape:
            raise ValueError("history must be x [T][N][nx], w / we [T][N]")
        if pf.nu > 0:
            u = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(T, pf.nu))
            up = u.ctypes.data_as(dp)
        else:
            up = C.cast(None, dp)
        epoch = pf._epoch if epoch is None else int(epoch)
        xb = np.zeros((T, int(M), pf.nx))
        check(pf._lib, pf._lib.llpf_smooth_history(pf._h, T, up, xf.ctypes.data_as(dp), wf.ctypes.data_as(dp),
                                                   wef.ctypes.data_as(dp), int(M), epoch, xb.ctypes.data_as(dp)))
        return xb, ll
    raise TypeError("smooth(pf, M, u, y) or smooth(pf, xf, wf, wef, ll, M, u, y)")


def last_smooth_ms(pf):
    ms = C.c_float()
    check(pf._lib, pf._lib.llpf_last_smooth_ms(pf._h, C.byref(ms)))
    return ms.value


def smoothed_mean(xb):
    """smoothed_mean(xb)  smoothing.jl:356-361 -> nx x T"""
    return np.asarray(xb).mean(axis=1).T


def smoothed_cov(xb):
    """smoothed_cov(xb)  smoothing.jl:368-372 -> list of T (nx x nx) sample covariances over the M trajectories"""
    xb = np.asarray(xb)
    return [np.atleast_2d(np.cov(xb[t].T)) for t in range(xb.shape[0])]


def smoothed_trajs(xb):
    """smoothed_trajs(xb)  smoothing.jl:379-383 -> (nx, M, T) array"""
    return np.ascontiguousarray(np.transpose(np.asarray(xb), (2, 1, 0)))


# ---------------------------------------------------------------------------------------------
# accessors (PFtypes.jl:296-334)
# ---------------------------------------------------------------------------------------------
def num_particles(pf):
    n = C.c_int64()
    check(pf._lib, pf._lib.llpf_num_particles(pf._h, C.byref(n)))
    return n.value


def index(pf):
    t = C.c_int64()
    check(pf._lib, pf._lib.llpf_index(pf._h, C.byref(t)))
    return t.value


def _get(pf, fn, shape, dtype=np.float64):
    a = np.zeros(shape, dtype=dtype)
    ptr = a.ctypes.data_as(dp if dtype == np.float64 else _abi.c_int64_p)
    check(pf._lib, fn(pf._h, ptr))
    return a


def particles(pf):
    return _get(pf, pf._lib.llpf_get_particles, (pf.N, pf.nx))


def weights(pf):
    return _get(pf, pf._lib.llpf_get_weights, (pf.N,))


def expweights(pf):
    return _get(pf, pf._lib.llpf_get_expweights, (pf.N,))


def ancestors(pf):
    """state(pf).j (1-based)"""
    return _get(pf, pf._lib.llpf_get_ancestors, (pf.N,), np.int64)


def bins(pf):
    return _get(pf, pf._lib.llpf_get_bins, (pf.N,))


def xprev(pf):
    """state(pf).xprev — at every API boundary xprev == x (copyto!(xprev, x) closes predict!, filtering.jl:151)"""
    return _get(pf, pf._lib.llpf_get_xprev, (pf.N, pf.nx))


def state(pf):
    x = particles(pf)
    return dict(x=x, xprev=xprev(pf), w=weights(pf), we=expweights(pf), j=ancestors(pf), bins=bins(pf), t=index(pf))


def set_state(pf, x, w, t):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(pf.N, pf.nx))
    w = np.ascontiguousarray(np.asarray(w, dtype=np.float64).reshape(pf.N))
    check(pf._lib, pf._lib.llpf_set_state(pf._h, x.ctypes.data_as(dp), w.ctypes.data_as(dp), int(t)))


def effective_particles(pf_or_we):
    """effective_particles(pf) / effective_particles(we) = 1/sum(abs2, we)  resample.jl:1-2"""
    if isinstance(pf_or_we, AbstractParticleFilter):
        v = C.c_double()
        check(pf_or_we._lib, pf_or_we._lib.llpf_effective_particles(pf_or_we._h, C.byref(v)))
        return v.value
    we = np.asarray(pf_or_we, dtype=np.float64)
    return 1.0 / float(np.sum(we * we))


def shouldresample(pf):
    v = C.c_int32()
    check(pf._lib, pf._lib.llpf_shouldresample(pf._h, C.byref(v)))
    return bool(v.value)


def weighted_mean(pf):
    xh = np.zeros(pf.nx)
    check(pf._lib, pf._lib.llpf_weighted_mean(pf._h, xh.ctypes.data_as(dp)))
    return xh


# ---------------------------------------------------------------------------------------------
# stand-alone numerics at the reference's function boundaries
# ---------------------------------------------------------------------------------------------
def resample(strategy, we, u01, M=None, j0=None, scan_mode="fast", device=0, return_bins=False):
    """resample(T, we, j, bins, M)  resample.jl:12-61 with the rand() draws supplied by the caller:
    systematic: u01 is the scalar rand() of :23 ; stratified: u01[M] are the rand() of :49.
    Returns 1-based Int64 indices (entries the reference leaves untouched keep j0, default 1:M)."""
    lib = _abi.load_library()
    we = np.ascontiguousarray(np.asarray(we, dtype=np.float64).reshape(-1))
    N = we.size
    M = N if M is None else int(M)
    j = np.arange(1, M + 1, dtype=np.int64) if j0 is None else np.array(j0, dtype=np.int64).copy()
    b = np.zeros(N)
    mode = SCAN_SERIAL if scan_mode in (SCAN_SERIAL, "serial") else SCAN_FAST
    if strategy is ResampleSystematic:
        check(lib, lib.llpf_resample_systematic(N, we.ctypes.data_as(dp), float(u01), M,
                                                 j.ctypes.data_as(_abi.c_int64_p), b.ctypes.data_as(dp), mode, device))
    elif strategy is ResampleStratified:
        u = np.ascontiguousarray(np.asarray(u01, dtype=np.float64).reshape(-1))
        if u.size != M:
            raise ValueError("stratified resampling needs M uniforms")
        check(lib, lib.llpf_resample_stratified(N, we.ctypes.data_as(dp), u.ctypes.data_as(dp), M,
                                                 j.ctypes.data_as(_abi.c_int64_p), b.ctypes.data_as(dp), mode, device))
    elif strategy is ResampleResidual:   # resample.jl:63-117 ; u01[M]: the rand() of :106 in draw order
        u = np.ascontiguousarray(np.asarray(u01, dtype=np.float64).reshape(-1))
        if u.size != M:
            raise ValueError("residual resampling needs M uniforms (only the first M - num are consumed)")
        check(lib, lib.llpf_resample_residual(N, we.ctypes.data_as(dp), u.ctypes.data_as(dp), M,
                                               j.ctypes.data_as(_abi.c_int64_p), b.ctypes.data_as(dp), mode, device))
    elif strategy is ResampleMetropolis:   # extension: u01 is the integer seed of the counter streams, B = 32 proposals
        check(lib, lib.llpf_resample_metropolis(N, we.ctypes.data_as(dp), M, 32, int(u01), j.ctypes.data_as(_abi.c_int64_p),
                                                 device))
    else:
        raise TypeError("strategy must be ResampleSystematic, ResampleStratified, ResampleResidual or ResampleMetropolis")
    return (j, b) if return_bins else j


def logsumexp(w, device=0):
    """ll = logsumexp!(w, we)  utils.jl:18-27 -> (ll, w_normalised, we)"""
    lib = _abi.load_library()
    w = np.ascontiguousarray(np.asarray(w, dtype=np.float64).reshape(-1)).copy()
    we = np.zeros_like(w)
    ll = C.c_double()
    check(lib, lib.llpf_logsumexp(w.size, w.ctypes.data_as(dp), we.ctypes.data_as(dp), C.byref(ll), device))
    return ll.value, w, we


def shard_blob(pf):
    """Opaque IPC descriptor of this rank's device arena (bytes) — exchange with every other rank."""
    n = C.c_size_t()
    check(pf._lib, pf._lib.llpf_shard_blob_size(C.byref(n)))
    buf = C.create_string_buffer(n.value)
    check(pf._lib, pf._lib.llpf_shard_export(pf._h, buf))
    return buf.raw


def connect_shards(pf, allgather=None):
    """Wire the ranks of a sharded filter together.  `allgather(bytes) -> list[bytes]` (ordered by rank) is any
    host-side all-gather; by default torch.distributed.all_gather_object on the default process group (NCCL or
    gloo: it only carries the 200-byte IPC descriptors — the data path is peer memory over NVLink)."""
    if pf.world <= 1:
        return
    mine = shard_blob(pf)
    if allgather is None:
        import torch.distributed as dist
        blobs = [None] * pf.world
        dist.all_gather_object(blobs, mine)
    else:
        blobs = list(allgather(mine))
    if len(blobs) != pf.world:
        raise ValueError("need one blob per rank")
    joined = b"".join(blobs)
    check(pf._lib, pf._lib.llpf_shard_connect(pf._h, C.create_string_buffer(joined, len(joined))))


def launch_count(pf):
    n = C.c_int64()
    check(pf._lib, pf._lib.llpf_launch_count(pf._h, C.byref(n)))
    return n.value


def last_run_ms(pf):
    v = C.c_float()
    check(pf._lib, pf._lib.llpf_last_run_ms(pf._h, C.byref(v)))
    return v.value
