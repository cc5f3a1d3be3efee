"""Rao-Blackwellized particle filter (reference src/rbpf.jl, test/test_rbpf.jl).

CPU: the restatement oracle/rbpf_ref.py against the closed form the reference's own tests use (`solkf.ll ≈ solrb.ll
rtol=1e-2`, test_rbpf.jl:111,141) and an offline NVRTC compile of the generated device source.
GPU: the device filter (llpf_b200.RBPF -> llpf_create_user with state hooks, csrc/llpf_rbpf.cuh) against that restatement
on identical counter-based RNG streams — ll, per-step ll, resample decisions, weighted means, final particles — for the
three model classes of test_rbpf.jl (mixed with An != 0, everything linear, everything nonlinear), plus the reference's
1 % criterion at its own size (N=500, T=500)."""
import math
import os
import sys

import numpy as np
import pytest

from oracle import rbpf_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def double_integrator(T, seed=0):
    """test_rbpf.jl:84-104: A=[1 .1;0 1], B=[0;1], C=[1 0], R1 = double_integrator_covariance(0.1) + 1e-6 I, R2 = 10"""
    rng = np.random.default_rng(seed)
    Ts = 0.1
    m = dict(A=[[1.0, 0.1], [0.0, 1.0]], B=[[0.0], [1.0]], C=[[1.0, 0.0]],
             R1=(np.array([[Ts ** 4 / 4, Ts ** 3 / 2], [Ts ** 3 / 2, Ts ** 2]]) + 1e-6 * np.eye(2)).tolist(), R2=[[10.0]],
             mu0=rng.standard_normal(2).tolist(), Sigma0=(2 * np.eye(2)).tolist())
    u = rng.standard_normal((T, 1))
    x = np.array(m["mu0"])
    L1 = np.linalg.cholesky(np.array(m["R1"]))
    y = np.zeros((T, 1))
    for t in range(T):
        y[t] = np.array(m["C"]) @ x + math.sqrt(10.0) * rng.standard_normal(1)
        x = np.array(m["A"]) @ x + np.array(m["B"]) @ u[t] + L1 @ rng.standard_normal(2)
    return m, u, y


def mixed_model(T, seed=1):
    """test_rbpf.jl:5-34: fn = xn, An = 0.5, A = 0.95, C = 1, g = xn, R1n = R1l = 0.01, R2 = 0.1, nu = 0"""
    rng = np.random.default_rng(seed)
    kf = dict(A=[[0.95]], B=[[]], C=[[1.0]], R1=[[0.01]], R2=[[0.1]], mu0=[1.0], Sigma0=[[1.0]])
    xn, xl = 1.0, 1.0
    y = np.zeros((T, 1))
    for t in range(T):
        y[t, 0] = xn + xl + math.sqrt(0.1) * rng.standard_normal()
        xn, xl = xn + 0.5 * xl + 0.1 * rng.standard_normal(), 0.95 * xl + 0.1 * rng.standard_normal()
    return kf, np.zeros((T, 0)), y


# ---------------------------------------------------------------------------------------------------------------------
# CPU
# ---------------------------------------------------------------------------------------------------------------------
def test_restatement_with_everything_linear_is_the_kalman_filter():
    """no random component at all (R1n = 0, d0n = 0): every particle IS the Kalman filter, so ll is the closed form"""
    m, u, y = double_integrator(120)
    kf = R.kalman_loglik(m["A"], m["B"], m["C"], m["R1"], m["R2"], m["mu0"], m["Sigma0"], u.tolist(), y.tolist())
    pf = R.RBPFRef(20, m, lambda xn, u, t: xn, lambda xn, u, t: [0.0], [[0.0]], ([0.0], [[0.0]]), An=None, seed=3)
    out = pf.run(u.tolist(), y.tolist())
    assert abs(out["ll"] - kf) <= 1e-9 * abs(kf)
    assert sum(out["resampled"]) == 0


def test_restatement_with_everything_nonlinear_meets_the_reference_criterion():
    """test_rbpf.jl:117-141: trivial inner Kalman filter (A=B=C=0), the whole model in fn / g: ll within 1 % of the KF"""
    m, u, y = double_integrator(150)
    kf = R.kalman_loglik(m["A"], m["B"], m["C"], m["R1"], m["R2"], m["mu0"], m["Sigma0"], u.tolist(), y.tolist())
    A, B, Cm = np.array(m["A"]), np.array(m["B"]), np.array(m["C"])
    kf2 = dict(A=[[0.0]], B=[[0.0]], C=[[0.0]], R1=[[1.0]], R2=m["R2"], mu0=[0.0], Sigma0=[[1.0]])
    pf = R.RBPFRef(300, kf2, lambda xn, u, t: (A @ np.array(xn) + B @ np.array(u)).tolist(),
                   lambda xn, u, t: (Cm @ np.array(xn)).tolist(), m["R1"], (m["mu0"], m["Sigma0"]), An=None, seed=4)
    out = pf.run(u.tolist(), y.tolist())
    assert abs(out["ll"] - kf) <= 1e-2 * abs(kf)
    assert sum(out["resampled"]) >= 1


def test_generated_device_source_compiles_offline():
    """NVRTC for sm_100a needs no GPU: the generated llpf_user source + csrc/llpf_rbpf.cuh + the engine compile cleanly"""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import nvrtc_probe as NP
    import llpf_b200 as L
    src = L.rbpf_source(1, 1, 1, 0, A=[[0.95]], B=None, C=[[1.0]], An=[[0.5]], R1l=[[0.01]], R1n=[[0.01]], R2=[[0.1]],
                        fn_body="fn[0] = xn[0];", g_body="yn[0] = xn[0];")
    assert "LLPF_USER_STATE_HOOKS" in src
    full = ("#define LLPF_USER_MODEL\n#define LLPF_USER_STATE_HOOKS\n#include \"llpf_engine.cuh\"\n" + src +
            "\nnamespace llpf {\ntemplate __global__ void k_engine<3, 1, 2, 0>(const __grid_constant__ EngineP, "
            "const __grid_constant__ ModelP<3, 1>);\n}\n")
    rc, _, log = NP.compile_engine(full.encode())[:3]
    assert rc == 0, log


def test_source_generator_flags_dimensions_and_literals():
    import llpf_b200 as L
    src = L.rbpf_source(2, 1, 1, 1, A=[[0.0]], B=[[0.3]], C=[[0.0]], An=None, R1l=[[1.0]], R1n=[[0.1, 0.02], [0.02, 0.3]],
                        R2=[[10.0]], fn_body="fn[0] = xn[0]; fn[1] = xn[1];", g_body="yn[0] = xn[0];")
    assert "llpf_rbpf::Consts<2, 1, 1, 1>" in src and "dynamics<4>" in src and "correct_state<4>" in src   # 2 + 1 + 1 components
    consts = src.split("const RB k = {")[1].split("}; return k;")[0]
    assert consts.rstrip().endswith(", 1, 1")                       # iszero(An) and iszero(C)  (rbpf.jl:179,247)
    assert float.fromhex("0x1.999999999999ap-4") == 0.1 and "0x1.999999999999ap-4" in consts   # exact hexadecimal literals
    src2 = L.rbpf_source(1, 2, 1, 0, A=[[1, 0.1], [0, 1]], B=None, C=[[1.0, 0]], An=[[0.5, 0]], R1l=[[1, 0], [0, 1]],
                         R1n=[[0.01]], R2=[[0.1]], fn_body="fn[0] = xn[0];", g_body="yn[0] = 0.0;")
    assert "dynamics<6>" in src2 and src2.split("const RB k = {")[1].split("}; return k;")[0].rstrip().endswith(", 0, 0")


def test_device_filters_fail_loudly_without_a_gpu(built):
    """no CPU fallback behind the new constructors either: without a device they raise, they do not compute on the host"""
    import ctypes as C
    import llpf_b200 as L
    n = C.c_int()
    if L.load_library().llpf_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is visible")
    kf, u, y, fn_c, g_c, fn_p, g_p, R1n, d0n, An = CASES["mixed"](3)
    with pytest.raises(L.LLPFError):
        _device_filter(L, 16, kf, fn_c, g_c, R1n, d0n, An, 0, 1)
    with pytest.raises(L.LLPFError):
        L.EnsembleKalmanFilter(L.LinearDynamics(np.eye(2), np.zeros((2, 1))), L.LinearMeasurement(np.eye(2)), np.eye(2), np.eye(2),
                               L.MvNormal(np.zeros(2), np.eye(2)), 16)


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
def _device_filter(L, N, kf, fn, g, R1n, d0n, An, nu, seed, thr=0.1, scan_mode="serial"):
    k = L.KalmanFilter(kf["A"], kf["B"] if nu else None, kf["C"], 0, kf["R1"], kf["R2"], L.MvNormal(np.array(kf["mu0"]), np.array(kf["Sigma0"])))
    mm = L.RBMeasurementModel(g, kf["R2"], 1)
    return L.RBPF(N, k, fn, mm, R1n, L.MvNormal(np.array(d0n[0]), np.array(d0n[1])), An=An, nu=nu, seed=seed,
                  resample_threshold=thr, scan_mode=scan_mode)


CASES = {
    # name: (model builder, fn device body, g device body, python fn, python g, R1n, d0n, An)
    "mixed": lambda T: (*mixed_model(T), "fn[0] = xn[0] + 0.05 * cos(t);", "yn[0] = xn[0];",
                        lambda xn, u, t: [xn[0] + 0.05 * math.cos(t)], lambda xn, u, t: [xn[0]],
                        [[0.01]], ([1.0], [[0.01]]), [[0.5]]),
    "linear": lambda T: (*double_integrator(T), "fn[0] = xn[0];", "yn[0] = 0.0;",
                         lambda xn, u, t: list(xn), lambda xn, u, t: [0.0], [[0.0]], ([0.0], [[0.0]]), None),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["mixed", "linear", "nonlinear"])
def test_device_rbpf_matches_restatement(gpu, case):
    L = gpu
    T, N, seed = 60, 400, 7
    if case == "nonlinear":
        m, u, y = double_integrator(T)
        A, B, Cm = np.array(m["A"]), np.array(m["B"]), np.array(m["C"])
        kf = dict(A=[[0.0]], B=[[0.0]], C=[[0.0]], R1=[[1.0]], R2=m["R2"], mu0=[0.0], Sigma0=[[1.0]])
        fn_c = "fn[0] = xn[0] + 0.1 * xn[1]; fn[1] = xn[1] + u[0];"
        g_c = "yn[0] = xn[0];"
        fn_p = lambda xn, u, t: (A @ np.array(xn) + B @ np.array(u)).tolist()   # noqa: E731
        g_p = lambda xn, u, t: (Cm @ np.array(xn)).tolist()                      # noqa: E731
        R1n, d0n, An = m["R1"], (m["mu0"], m["Sigma0"]), None
    else:
        kf, u, y, fn_c, g_c, fn_p, g_p, R1n, d0n, An = CASES[case](T)
    nu = u.shape[1]
    thr = 0.5 if case != "linear" else 0.1
    ref = R.RBPFRef(N, kf, fn_p, g_p, R1n, d0n, An=An, resample_threshold=thr, seed=seed)
    out = ref.run(u.tolist(), y.tolist(), epoch=2)
    pf = _device_filter(L, N, kf, fn_c, g_c, R1n, d0n, An, nu, seed, thr)
    sol = L.forward_trajectory(pf, u if nu else None, y, epoch=2)
    assert np.array_equal(sol.extra["resampled"], out["resampled"])
    assert np.allclose(sol.extra["ll_steps"], out["ll_steps"], rtol=1e-9, atol=1e-9)
    assert abs(sol.ll - out["ll"]) <= 1e-9 * max(1.0, abs(out["ll"]))
    nst = ref.nxn + ref.nxl
    assert np.allclose(sol.extra["xhat"][:, :nst], np.array(out["xhat"]), rtol=0, atol=1e-9)
    assert np.allclose(L.particles(pf), np.array(ref.particles_flat()), rtol=0, atol=1e-8)
    if case != "linear":
        assert sum(out["resampled"]) >= 2
    # the history holds the particles AFTER correct! (filtering.jl:357), i.e. with the Kalman update applied
    xn, xl, Rc = L.rb_particles(pf, sol.x[-1])
    assert xn.shape == (N, ref.nxn) and xl.shape == (N, ref.nxl) and Rc.shape == (N, ref.nxl, ref.nxl)
    assert np.all(np.linalg.eigvalsh(Rc) > -1e-12)


@pytest.mark.gpu
def test_device_rbpf_step_verbs_equal_the_fused_trajectory(gpu):
    L = gpu
    kf, u, y, fn_c, g_c, fn_p, g_p, R1n, d0n, An = CASES["mixed"](30)
    pf = _device_filter(L, 256, kf, fn_c, g_c, R1n, d0n, An, 0, 5, thr=0.5)
    sol = L.forward_trajectory(pf, None, y, epoch=1)
    fused = L.particles(pf).copy()
    L.reset(pf, epoch=1)
    ll = 0.0
    for t in range(len(y)):
        ll += L.correct(pf, None, y[t], None, t * pf.Ts)[0]
        L.predict(pf, None, None, t * pf.Ts)
    assert abs(ll - sol.ll) <= 1e-12 * abs(sol.ll)
    assert np.array_equal(L.particles(pf), fused)


@pytest.mark.gpu
def test_device_rbpf_refuses_the_auxiliary_filter(gpu):
    """the auxiliary filter's two-stage step does not call the state hooks: it must fail loudly, not run a wrong filter"""
    L = gpu
    kf, u, y, fn_c, g_c, fn_p, g_p, R1n, d0n, An = CASES["mixed"](5)
    pf = _device_filter(L, 64, kf, fn_c, g_c, R1n, d0n, An, 0, 1)
    with pytest.raises(L.LLPFError):
        L.AuxiliaryParticleFilter(pf)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["linear", "nonlinear"])
def test_device_rbpf_meets_the_reference_criterion(gpu, case):
    """test_rbpf.jl:111,141 at the reference's own size: N = 500 particles, T = 500 steps, ll within 1 % of the KF"""
    L = gpu
    m, u, y = double_integrator(500, seed=2)
    kfll = R.kalman_loglik(m["A"], m["B"], m["C"], m["R1"], m["R2"], m["mu0"], m["Sigma0"], u.tolist(), y.tolist())
    if case == "linear":
        pf = _device_filter(L, 500, m, "fn[0] = xn[0];", "yn[0] = 0.0;", [[0.0]], ([0.0], [[0.0]]), None, 1, 1, scan_mode="fast")
    else:
        kf = dict(A=[[0.0]], B=[[0.0]], C=[[0.0]], R1=[[1.0]], R2=m["R2"], mu0=[0.0], Sigma0=[[1.0]])
        pf = _device_filter(L, 500, kf, "fn[0] = xn[0] + 0.1 * xn[1]; fn[1] = xn[1] + u[0];", "yn[0] = xn[0];", m["R1"],
                            (m["mu0"], m["Sigma0"]), None, 1, 1, scan_mode="fast")
    sol = L.forward_trajectory(pf, u, y, history=False)
    assert abs(sol.ll - kfll) <= 1e-2 * abs(kfll), (sol.ll, kfll)
    if case == "linear":
        assert abs(sol.ll - kfll) <= 1e-9 * abs(kfll)     # exact: every particle is the Kalman filter
