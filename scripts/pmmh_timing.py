"""PMMH on the device: the reference's example (example_lineargaussian.jl:195-223: 1200 loglik evaluations of a filter
with N=1000, T=200 take "about half a minute") on one B200 — one chain, then 16 / 64 concurrent chains on host threads."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llpf_b200 as L
from models import lg_model

s = lg_model(2, 2, 2, seed=0)
T, N = 200, 1000
u = np.random.default_rng(0).standard_normal((T, 2))
gen = s.oracle_filter(16, seed=1)
_, y = gen.simulate(u, 7)

def ffp(theta, pf=None):
    d1, d2 = L.MvNormal(math.exp(theta[0]) * np.eye(2)), L.MvNormal(math.exp(theta[1]) * np.eye(2))
    if pf is None:
        return L.ParticleFilter(N, L.LinearDynamics(s.A, s.B), L.LinearMeasurement(s.C), d1, d2, L.MvNormal(s.mu0, s.Sigma0), seed=4)
    return L.set_model(pf, dynamics_density=d1, measurement_density=d2)

priors = [L.Normal(0, 1.0), L.Normal(0, 1.0)]
rng = np.random.default_rng(3)
ll = L.log_likelihood_fun(ffp, priors, u, y)
ll(np.zeros(2))
R = 1200
t0 = time.perf_counter()
p, l = L.metropolis(ll, R, np.array([0.5, -0.5]), lambda th: th + 0.1 * rng.standard_normal(2), rng)
dt = time.perf_counter() - t0
print(f"1 chain: {R} loglik evaluations (N={N}, T={T}) in {dt:.3f} s = {R * N * T / dt / 1e6:.1f} M particle-steps/s; "
      f"posterior mean {p[300:].mean(axis=0)}; kernel {L.last_run_ms(ll.state['pf']):.3f} ms per evaluation", flush=True)
for K in (16, 64):
    t0 = time.perf_counter()
    out = L.metropolis_threaded(0, lambda: L.log_likelihood_fun(ffp, priors, u, y), 300, np.array([0.5, -0.5]), None, nthreads=K, seed=2)
    dt = time.perf_counter() - t0
    print(f"{K} chains x 300 evaluations in {dt:.3f} s = {K * 300 * N * T / dt / 1e6:.1f} M particle-steps/s "
          f"({K * 300 / dt:.0f} loglik/s)", flush=True)

# batched: one launch per MCMC iteration for all chains (llpf_run_batch, one thread block per chain)
def ffp1(theta, pf=None):
    d1, d2 = L.MvNormal(math.exp(theta[0]) * np.eye(2)), L.MvNormal(math.exp(theta[1]) * np.eye(2))
    if pf is None:
        return L.ParticleFilter(N, L.LinearDynamics(s.A, s.B), L.LinearMeasurement(s.C), d1, d2, L.MvNormal(s.mu0, s.Sigma0), seed=4, single_block=True)
    return L.set_model(pf, dynamics_density=d1, measurement_density=d2)
for K in (64, 296, 592, 1184):
    pfs = [ffp1(np.zeros(2)) for _ in range(K)]
    L.loglik_batch(pfs, u, y)
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps):
        L.loglik_batch(pfs, u, y)
    dt = (time.perf_counter() - t0) / reps
    print(f"loglik_batch {K} chains: {dt * 1e3:.3f} ms per launch (kernel {L.last_run_ms(pfs[0]):.3f} ms) = {K / dt:.0f} loglik/s = "
          f"{K * N * T / dt / 1e9:.2f} G particle-steps/s", flush=True)
    del pfs
def draw(rng):
    return lambda th: th + 0.1 * rng.standard_normal(2)
draw.wants_rng = True
for K in (64, 296):
    t0 = time.perf_counter()
    out = L.metropolis_batched(0, ffp1, priors, u, y, 100, np.array([0.5, -0.5]), draw, nchains=K, seed=2)
    dt = time.perf_counter() - t0
    print(f"metropolis_batched {K} chains x 100 iterations in {dt:.3f} s = {K * 100 / dt:.0f} loglik/s (host proposal / set_model included)", flush=True)
