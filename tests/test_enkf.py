"""Ensemble Kalman filter (reference src/enkf.jl, test/test_enkf.jl).

CPU: the restatement oracle/enkf_ref.py against the closed-form Kalman filter with the reference's own criteria
(test_enkf.jl:115-119: sse_enkf < 1.2 sse_kf, |ll_enkf - ll_kf| < 5 at N = 500, T = 200 on the linear test system).
GPU: the device filter (llpf_enkf_* entry points, csrc/llpf_enkf.cuh) against the restatement on identical counter-based
RNG streams — every quantity of the KalmanFilteringSolution to 1e-9 (the reference sums the ensemble sequentially, the
device with a deterministic tree: agreement to rounding, stated) — the step verbs against the fused trajectory, inflation,
the quadtank dynamics, and the reference's criteria at the reference's size."""
import math

import numpy as np
import pytest

from oracle import enkf_ref as E
from oracle import rbpf_ref as R

A = [[0.99, 0.1], [0.0, 0.2]]          # test_enkf.jl:22-24
B = [[-0.74, 1.61], [-1.44, 1.75]]
Cm = [[1.0, 0.0], [0.0, 1.0]]
R1 = [[1.0, 0.0], [0.0, 1.0]]
R2 = [[1.0, 0.0], [0.0, 1.0]]


def _system(T, seed=0):
    rng = np.random.default_rng(seed)
    mu0 = rng.standard_normal(2)
    S0 = 4.0 * np.eye(2)
    u = rng.standard_normal((T, 2))
    x = mu0 + 2.0 * rng.standard_normal(2)
    xs, y = np.zeros((T, 2)), np.zeros((T, 2))
    for t in range(T):
        xs[t] = x
        y[t] = np.array(Cm) @ x + rng.standard_normal(2)
        x = np.array(A) @ x + np.array(B) @ u[t] + rng.standard_normal(2)
    return mu0.tolist(), S0.tolist(), u, y, xs


def _lin(x, u, t):
    return [A[0][0] * x[0] + A[0][1] * x[1] + (B[0][0] * u[0] + B[0][1] * u[1]),
            A[1][0] * x[0] + A[1][1] * x[1] + (B[1][0] * u[0] + B[1][1] * u[1])]


def _kf_filtered_means(mu0, S0, u, y):
    """xt of forward_trajectory(kf, u, y): numpy Kalman filter (filtering.jl:52-133)"""
    x, P = np.array(mu0), np.array(S0)
    An, Bn, Cn = np.array(A), np.array(B), np.array(Cm)
    out = np.zeros((len(y), 2))
    for t in range(len(y)):
        S = Cn @ P @ Cn.T + np.array(R2)
        K = P @ Cn.T @ np.linalg.inv(S)
        x = x + K @ (y[t] - Cn @ x)
        P = (np.eye(2) - K @ Cn) @ P
        out[t] = x
        x = An @ x + Bn @ u[t]
        P = An @ P @ An.T + np.array(R1)
    return out


def test_restatement_meets_the_reference_criteria_on_the_linear_system():
    mu0, S0, u, y, xs = _system(200, seed=1)
    kfll = R.kalman_loglik(A, B, Cm, R1, R2, mu0, S0, u.tolist(), y.tolist())
    ref = E.EnKFRef(_lin, Cm, R1, R2, mu0, S0, 500, seed=3)
    out = ref.forward_trajectory(u.tolist(), y.tolist(), epoch=1)
    assert abs(out["ll"] - kfll) < 5.0                                       # test_enkf.jl:119
    sse = lambda a: float(np.sum((xs - np.array(a)) ** 2))                   # noqa: E731
    assert sse(out["xt"]) < 1.2 * sse(_kf_filtered_means(mu0, S0, u, y))     # :115
    assert ref.t == 200 and len(out["x"]) == 200


def test_restatement_statistics_and_inflation():
    mu0, S0, u, y, _ = _system(5, seed=2)
    ref = E.EnKFRef(_lin, Cm, R1, R2, mu0, S0, 400, seed=5)
    X = np.array(ref.X)
    assert np.allclose(ref.x, X.mean(axis=0), atol=1e-12) and np.allclose(ref.R, np.cov(X.T), atol=1e-12)
    assert np.linalg.norm(np.array(ref.x) - np.array(mu0)) < 1.0             # test_enkf.jl:46
    plain = E.EnKFRef(_lin, Cm, R1, R2, mu0, S0, 400, seed=5)
    infl = E.EnKFRef(_lin, Cm, R1, R2, mu0, S0, 400, seed=5, inflation=1.05)
    plain.predict(u[0].tolist(), 0.0); infl.predict(u[0].tolist(), 0.0)
    assert np.allclose(infl.x, plain.x, atol=1e-12)                          # the spread grows, the mean stays
    assert np.allclose(infl.R, 1.05 ** 2 * np.array(plain.R), rtol=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
def _device(L, N, mu0, S0, seed, **kw):
    return L.EnsembleKalmanFilter(L.LinearDynamics(np.array(A), np.array(B)), L.LinearMeasurement(np.array(Cm)), np.array(R1),
                                  np.array(R2), L.MvNormal(np.array(mu0), np.array(S0)), N, nu=2, ny=2, seed=seed, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("inflation", [1.0, 1.05])
def test_device_enkf_matches_restatement(gpu, inflation):
    L = gpu
    T, N, seed = 40, 700, 9
    mu0, S0, u, y, _ = _system(T, seed=4)
    ref = E.EnKFRef(_lin, Cm, R1, R2, mu0, S0, N, seed=seed, inflation=inflation)
    out = ref.forward_trajectory(u.tolist(), y.tolist(), epoch=2)
    enkf = _device(L, N, mu0, S0, seed, inflation=inflation)
    sol = L.enkf_forward_trajectory(enkf, u, y, epoch=2)
    tol = dict(rtol=1e-9, atol=1e-9)
    assert np.allclose(sol.extra["ll_steps"], out["ll_steps"], **tol)
    assert abs(sol.ll - out["ll"]) <= 1e-9 * abs(out["ll"])
    for key in ("x", "xt", "R", "Rt", "e", "S", "K"):
        assert np.allclose(getattr(sol, key), np.array(out[key]), **tol), key
    assert np.allclose(L.enkf_particles(enkf), np.array(ref.X), **tol)
    assert np.allclose(L.enkf_state(enkf), ref.x, **tol) and np.allclose(L.enkf_covariance(enkf), ref.R, **tol)
    assert enkf.t == T


@pytest.mark.gpu
def test_device_enkf_step_verbs_equal_the_fused_trajectory(gpu):
    L = gpu
    T, N = 12, 300
    mu0, S0, u, y, _ = _system(T, seed=6)
    enkf = _device(L, N, mu0, S0, 2)
    sol = L.enkf_forward_trajectory(enkf, u, y, epoch=3)
    fused = L.enkf_particles(enkf).copy()
    L.enkf_reset(enkf, epoch=3)
    assert enkf.t == 0 and L.enkf_state(enkf).shape == (2,) and L.enkf_covariance(enkf).shape == (2, 2)   # test_enkf.jl:38-41,51
    ll = 0.0
    for k in range(T):
        r = L.enkf_update(enkf, u[k], y[k], None, k * enkf.Ts)
        assert set(r) == {"ll", "e", "S", "K"}                                                           # :74-77
        assert np.allclose(r["e"], sol.e[k], rtol=0, atol=1e-12)
        ll += r["ll"]
    assert enkf.t == T
    assert abs(ll - sol.ll) <= 1e-12 * abs(sol.ll)
    assert np.array_equal(L.enkf_particles(enkf), fused)


@pytest.mark.gpu
def test_device_enkf_meets_the_reference_criteria(gpu):
    """test_enkf.jl:98-119 at the reference's size (N = 500, T = 200), and with an ensemble only a GPU would use"""
    L = gpu
    mu0, S0, u, y, xs = _system(200, seed=1)
    kfll = R.kalman_loglik(A, B, Cm, R1, R2, mu0, S0, u.tolist(), y.tolist())
    sse_kf = float(np.sum((xs - _kf_filtered_means(mu0, S0, u, y)) ** 2))
    for N, tol_ll in ((500, 5.0), (1 << 18, 0.5)):
        sol = L.enkf_forward_trajectory(_device(L, N, mu0, S0, 3), u, y)
        assert abs(sol.ll - kfll) < tol_ll, (N, sol.ll, kfll)
        assert float(np.sum((xs - sol.xt) ** 2)) < 1.2 * sse_kf


@pytest.mark.gpu
def test_device_enkf_quadtank(gpu):
    """nonlinear dynamics through the engine's quadtank descriptor (example_quadtank.jl:91-106): runs, finite, tracks levels"""
    L = gpu
    from llpf_b200 import workloads as W
    spec = W.QuadtankSpec()
    T = 200
    u = spec.inputs(T)
    x_true, y = W.simulate_quadtank(spec, u, seed=0)
    enkf = L.EnsembleKalmanFilter(spec.dynamics(), L.LinearMeasurement(spec.C), spec.R1, spec.R2,
                                  L.MvNormal(spec.x0, spec.R1), 2000, seed=1)
    sol = L.enkf_forward_trajectory(enkf, u, y)
    assert math.isfinite(sol.ll)
    assert np.sqrt(np.mean((sol.xt[:, :2] - x_true[:, :2]) ** 2)) < 0.2
