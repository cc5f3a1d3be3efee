"""EnKF on the device: time per forward_trajectory step for EnKF-sized and GPU-sized ensembles (4-state LG model of config 2),
and the RBPF (mixed model of test_rbpf.jl) for comparison with the plain particle filter."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llpf_b200 as L
from llpf_b200 import workloads as W

spec = W.lg_spec(4, 2, 2, seed=0)
T = 200
u = np.random.default_rng(0).standard_normal((T, 2)); _, y = W.simulate_lg(spec, u, seed=1)
for log2n in (9, 14, 17, 20):
    N = 1 << log2n
    e = L.EnsembleKalmanFilter(L.LinearDynamics(spec.A, spec.B), L.LinearMeasurement(spec.C), spec.R1, spec.R2,
                               L.MvNormal(spec.mu0, spec.Sigma0), N, seed=1)
    best = 1e9
    for r in range(3):
        sol = L.enkf_forward_trajectory(e, u, y)
        ms = np.zeros(1, dtype=np.float32)
        import ctypes as C
        e._lib.llpf_last_run_ms(e._h, ms.ctypes.data_as(C.POINTER(C.c_float)))
        best = min(best, float(ms[0]))
    print(f"EnKF N=2^{log2n} T={T}: {best:8.3f} ms  {best / T * 1e3:7.2f} us/step  {N * T / best / 1e6:8.2f} G member-steps/s  "
          f"ll={sol.ll:.3f}", flush=True)

# RBPF: mixed model (An != 0), composite particle of 3 doubles
kf = L.KalmanFilter([[0.95]], None, [[1.0]], 0, [[0.01]], [[0.1]], L.MvNormal(np.array([1.0]), np.array([[1.0]])))
rng = np.random.default_rng(1)
yy = 2.0 + 0.3 * rng.standard_normal((T, 1))
for log2n in (10, 20):
    N = 1 << log2n
    pf = L.RBPF(N, kf, "fn[0] = xn[0];", L.RBMeasurementModel("yn[0] = xn[0];", [[0.1]], 1), [[0.01]],
                L.MvNormal(np.array([1.0]), np.array([[0.01]])), An=[[0.5]], nu=0, seed=1, resample_threshold=0.5)
    best = 1e9
    for r in range(3):
        d = L.loglik(pf, None, yy, epoch=r + 1, details=True)
        best = min(best, L.last_run_ms(pf))
    print(f"RBPF N=2^{log2n} T={T}: {best:8.3f} ms  {best / T * 1e3:7.2f} us/step  {N * T / best / 1e6:8.2f} G particle-steps/s  "
          f"rho={d['resampled'].mean():.2f} ll={d['ll']:.3f}", flush=True)
