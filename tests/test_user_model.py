"""User-defined dynamics / measurement functions (the reference's closures, PFtypes.jl:116,128,232,255) as CUDA device code
compiled at run time into the fused sweep (llpf_create_user): GPU vs the CPU oracle evaluating the SAME functions as host
callbacks (OracleFilter.set_user_functions) on identical counter-based RNG streams."""
import math

import numpy as np
import pytest

from models import lg_model
from oracle import oracle as O

pytestmark = pytest.mark.gpu

DYN = """
const double x0 = x[0] + p[0] * sin(x[1]) + u[0];
const double x1 = p[1] * x[1] + 0.1 * cos(t);
x[0] = x0; x[1] = x1;
"""
LIK = """
const double r = y[0] - x[0] * x[0];
return -0.5 * r * r / p[2] - 0.5 * log(6.283185307179586 * p[2]);
"""
MEAS = "yh[0] = x[0] * x[0];"


def _py_dyn(p):
    return lambda x, u, t: np.array([x[0] + p[0] * math.sin(x[1]) + u[0], p[1] * x[1] + 0.1 * math.cos(t)])


def _py_lik(p):
    return lambda x, u, y, t: -0.5 * (y[0] - x[0] * x[0]) ** 2 / p[2] - 0.5 * math.log(6.283185307179586 * p[2])


def _oracle(N, filt, p, seed, thr):
    R1 = np.diag([0.05, 0.02])
    om = O.ModelArrays(2, 1, 1, np.array([[1.0, 0.0]]), R1, np.array([[p[2]]]), np.array([0.5, -0.2]), 0.25 * np.eye(2),
                       A=np.eye(2), B=np.zeros((2, 1)), dynamics=0)
    of = O.OracleFilter(om, N, filter=filt, resample_threshold=thr, seed=seed)
    of.set_user_functions(dynamics=_py_dyn(p), loglik=_py_lik(p))
    return of, R1


def _data(T, p, seed=3):
    rng = np.random.default_rng(seed)
    u = 0.3 * rng.standard_normal((T, 1))
    x = np.array([0.5, -0.2])
    y = np.zeros((T, 1))
    f = _py_dyn(p)
    for t in range(T):
        y[t, 0] = x[0] ** 2 + math.sqrt(p[2]) * rng.standard_normal()
        x = f(x, u[t], float(t)) + np.sqrt([0.05, 0.02]) * rng.standard_normal(2)
    return u, y


@pytest.mark.parametrize("kind", ["advanced", "pf", "aux", "aux_advanced"])
def test_nonlinear_user_model_matches_oracle(gpu, kind):
    L = gpu
    p = np.array([0.5, 0.9, 0.3])
    N, T, thr, seed = 3000, 40, 0.5, 11
    u, y = _data(T, p)
    filt = {"pf": 0, "advanced": 1, "aux": 2, "aux_advanced": 3}[kind]
    of, R1 = _oracle(N, filt, p, seed, thr)
    d0 = L.MvNormal(np.array([0.5, -0.2]), 0.25 * np.eye(2))
    if kind in ("pf", "aux"):
        pf = L.ParticleFilter(N, L.CudaDynamics(DYN, nu=1), L.CudaMeasurement(MEAS), L.MvNormal(np.zeros(2), R1),
                              L.MvNormal(np.zeros(1), np.array([[p[2]]])), d0, p=p, seed=seed, resample_threshold=thr,
                              scan_mode="serial")
    else:
        pf = L.AdvancedParticleFilter(N, L.CudaDynamics(DYN, nu=1), None, L.CudaLikelihood(LIK, ny=1),
                                      L.MvNormal(np.zeros(2), R1), d0, p=p, seed=seed, resample_threshold=thr,
                                      scan_mode="serial")
    if kind.startswith("aux"):
        pf = L.AuxiliaryParticleFilter(pf)
    for conv in ("ft", "loglik"):
        if conv == "ft":
            sol = L.forward_trajectory(pf, u, y, epoch=2)
            ref = of.forward_trajectory(u, y, epoch=2, history=True)
            got_ll, res = sol.ll, sol.extra["resampled"]
            assert np.allclose(sol.x, ref["x"], rtol=0, atol=1e-9)
            assert np.allclose(sol.w, ref["w"], rtol=0, atol=1e-8)
        else:
            r = L.loglik(pf, u, y, epoch=3, details=True)
            ref = of.loglik(u, y, epoch=3)
            got_ll, res = r["ll"], r["resampled"]
        assert np.array_equal(res, ref["resampled"])
        assert abs(got_ll - ref["ll"]) <= 1e-9 * max(1.0, abs(ref["ll"])), (kind, conv, got_ll, ref["ll"])
        assert np.array_equal(L.ancestors(pf), of.ancestors)
        assert np.allclose(L.particles(pf), of.particles, rtol=0, atol=1e-9)
    assert ref["resampled"].sum() >= 3


def test_user_model_any_dimension_and_parameter_override(gpu):
    """nx = 5 / ny = 3 has no pre-compiled engine: as a user model the instantiation is generated at run time.  The
    per-call parameter vector `p` (filtering.jl:140,164) replaces the one given at construction without recompiling."""
    L = gpu
    from llpf_b200 import filters as F
    s = lg_model(5, 2, 3, seed=4)
    N, T = 2000, 30
    u = np.random.default_rng(1).standard_normal((T, 2))
    _, y = s.oracle_filter(16, seed=1).simulate(u, 7)
    # dynamics x+ = p[0] * (A x + B u): p[0] = 1 reproduces the descriptor model
    body = F._linear_dynamics_body(s.A, s.B, 5, 2) + "\nfor (int k = 0; k < 5; ++k) x[k] *= p[0];"
    pf = L.ParticleFilter(N, L.CudaDynamics(body, nu=2), L.LinearMeasurement(s.C), L.MvNormal(np.zeros(5), s.R1),
                          L.MvNormal(np.zeros(3), s.R2), L.MvNormal(s.mu0, s.Sigma0), p=[1.0], seed=5, scan_mode="serial")
    ref = s.oracle_filter(N, seed=5).loglik(u, y, epoch=1)
    got = L.loglik(pf, u, y, epoch=1, details=True)
    assert np.array_equal(got["resampled"], ref["resampled"])
    assert abs(got["ll"] - ref["ll"]) <= 1e-9 * abs(ref["ll"])
    a = L.loglik(pf, u, y, [0.9], epoch=1)          # per-call override
    pf2 = L.ParticleFilter(N, L.CudaDynamics(body, nu=2), L.LinearMeasurement(s.C), L.MvNormal(np.zeros(5), s.R1),
                           L.MvNormal(np.zeros(3), s.R2), L.MvNormal(s.mu0, s.Sigma0), p=[0.9], seed=5, scan_mode="serial")
    assert a == L.loglik(pf2, u, y, epoch=1) and a != got["ll"]
    # descriptor models carry no parameter vector: an override must not be silently ignored
    s4 = lg_model(4, 2, 2, seed=0)
    with pytest.raises(TypeError):
        L.loglik(s4.particle_filter(64, seed=1), u[:, :2], y[:, :2], [1.0])
    # ... and nx = 5 exists as a descriptor model only through the run-time path
    with pytest.raises(L.LLPFError):
        s.particle_filter(64, seed=1)


def test_user_model_compile_error_is_reported(gpu):
    L = gpu
    with pytest.raises(L.LLPFError) as e:
        L.AdvancedParticleFilter(64, L.CudaDynamics("x[0] = undefined_symbol;", nu=0), None, L.CudaLikelihood("return 0.0;", ny=1),
                                 L.MvNormal(np.zeros(1), np.eye(1)), L.MvNormal(np.zeros(1), np.eye(1)))
    assert "does not compile" in str(e.value) and "undefined_symbol" in str(e.value)
