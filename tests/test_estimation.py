"""Host-side logic around the hot path: PMMH driver (smoothing.jl:266-347) and weighted statistics
(filtering.jl:570-595).  CPU tests use analytic log-likelihoods; the GPU tests drive the device `loglik`."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llpf_b200 as L  # noqa: E402
from models import lg_model  # noqa: E402


def test_weighted_quantile_and_cov_reduce_to_plain_statistics():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 101, 2))
    we = np.full((3, 101), 1 / 101)
    for q in (0.1, 0.5, 0.9):
        assert np.allclose(L.weighted_quantile(x, we, q), np.quantile(x, q, axis=1))
    cov = L.weighted_cov(x, we)
    for t in range(3):
        assert np.allclose(cov[t], np.cov(x[t].T))
    # zero-weight samples are ignored; a point mass gives that point
    w2 = np.zeros((1, 5)); w2[0, 3] = 1.0
    xs = np.arange(5.0).reshape(1, 5, 1)
    assert L.weighted_quantile(xs, w2, 0.3)[0, 0] == 3.0
    # integer-ratio weights == repeated samples (ProbabilityWeights interpolate on cumulative weight)
    v = np.array([1.0, 2.0, 4.0]).reshape(1, 3, 1)
    w = np.array([[0.25, 0.25, 0.5]])
    assert L.weighted_quantile(v, w, 0.0)[0, 0] == 1.0 and L.weighted_quantile(v, w, 1.0)[0, 0] == 4.0


def test_metropolis_samples_a_gaussian_target():
    rng = np.random.default_rng(1)
    target = lambda th: -0.5 * float(((th[0] - 1.5) / 0.3) ** 2)
    draw = lambda th: th + 0.3 * rng.standard_normal(1)
    p, l = L.metropolis(target, 6000, np.array([0.5]), draw, rng)
    assert p.shape == (6000, 1) and l.shape == (6000,)
    assert abs(p[1000:].mean() - 1.5) < 0.05 and abs(p[1000:].std() - 0.3) < 0.05
    assert np.all(l == np.array([target(t) for t in p]))            # the stored ll belongs to the stored sample
    with pytest.raises(ValueError):
        L.naive_sampler(np.array([0.0, 1.0]))


def test_log_likelihood_fun_prior_short_circuit_and_threads():
    calls = []

    class FakePF:
        pass

    def ffp(theta, pf=None):
        calls.append(pf)
        return FakePF()

    ll = L.log_likelihood_fun(ffp, [L.Uniform(0, 1)], None, None)
    assert ll(np.array([2.0])) == -math.inf and calls == []        # outside the prior: the filter is never built
    with pytest.raises(ValueError):
        ll(np.array([0.1, 0.2]))
    target = lambda th: -0.5 * float((th[0] / 0.5) ** 2)
    out = L.metropolis_threaded(100, target, 400, np.array([0.3]), None, nthreads=3, seed=5)
    assert out.shape == (3 * 300, 2)
    assert np.allclose(out[:, 1], -0.5 * (out[:, 0] / 0.5) ** 2)


# ---- on the device ------------------------------------------------------------------------------------------------
def _pmmh_problem(T=100, dims=(2, 1, 1)):
    s = lg_model(*dims, seed=3 if dims == (2, 1, 1) else 0)
    u = np.random.default_rng(0).standard_normal((T, dims[1]))
    gen = s.oracle_filter(16, seed=1)
    _, y = gen.simulate(u, 7)
    return s, u, y


@pytest.mark.gpu
def test_set_model_equals_fresh_filter(gpu):
    s, u, y = _pmmh_problem()
    pf = s.particle_filter(1000, seed=2)
    a = L.loglik(pf, u, y, epoch=1)
    L.set_model(pf, dynamics_density=L.MvNormal(0.25 * np.eye(2)), measurement_density=L.MvNormal(2.0 * np.eye(1)))
    b = L.loglik(pf, u, y, epoch=1)
    s2 = lg_model(2, 1, 1, seed=3, r1=0.25, r2=2.0)
    c = L.loglik(s2.particle_filter(1000, seed=2), u, y, epoch=1)
    assert b == c and a != b


@pytest.mark.gpu
def test_pmmh_on_device_finds_the_noise_level(gpu):
    """example_lineargaussian.jl:195-223 in miniature: θ = log of the dynamics / measurement noise variances, the data
    were simulated with variances (1, 1): the chain concentrates around θ = 0."""
    s, u, y = _pmmh_problem(T=200, dims=(2, 2, 2))
    Np = 2000

    def ffp(theta, pf=None):
        d1, d2 = L.MvNormal(math.exp(theta[0]) * np.eye(2)), L.MvNormal(math.exp(theta[1]) * np.eye(2))
        if pf is None:
            return L.ParticleFilter(Np, L.LinearDynamics(s.A, s.B), L.LinearMeasurement(s.C), d1, d2,
                                    L.MvNormal(s.mu0, s.Sigma0), seed=4)
        return L.set_model(pf, dynamics_density=d1, measurement_density=d2)

    priors = [L.Normal(0, 1.0), L.Normal(0, 1.0)]
    ll = L.log_likelihood_fun(ffp, priors, u, y)
    rng = np.random.default_rng(3)
    draw = lambda th: th + 0.1 * rng.standard_normal(2)
    p, l = L.metropolis(ll, 500, np.array([0.5, -0.5]), draw, rng)
    assert np.all(np.isfinite(l))
    post = p[150:]
    accept = np.mean(np.any(np.diff(p, axis=0) != 0, axis=1))
    assert accept > 0.03
    assert np.all(np.abs(post.mean(axis=0)) < 0.5)
    assert l[150:].mean() > l[0]
    # independent chains on their own handles / streams
    out = L.metropolis_threaded(50, lambda: L.log_likelihood_fun(ffp, priors, u, y), 120, np.array([0.3, -0.3]),
                                None, nthreads=4, seed=1)
    assert out.shape == (4 * 70, 3) and np.all(np.isfinite(out))
    assert np.all(np.abs(out[:, :2].mean(axis=0)) < 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("dims,aux", [((2, 2, 2), False), ((4, 2, 2), False), ((2, 1, 1), True)])
def test_batched_loglik_is_bit_identical_to_single_handles(gpu, dims, aux):
    """llpf_run_batch: C chains (own model, seed, epoch each) in ONE launch, one thread block per chain — every value
    bit-identical to the same chain evaluated on its own handle, and in parity with the CPU oracle."""
    s, u, y = _pmmh_problem(T=60, dims=dims)
    N, Cn = 1000, 37
    rng = np.random.default_rng(5)
    pfs, specs = [], []
    for c in range(Cn):
        sc = lg_model(*dims, seed=3 if dims == (2, 1, 1) else 0, r1=float(np.exp(0.3 * rng.standard_normal())),
                      r2=float(np.exp(0.3 * rng.standard_normal())))
        mk = sc.aux_filter if aux else sc.particle_filter
        pfs.append(mk(N, seed=100 + c, single_block=True))
        specs.append(sc)
    epochs = [7 + 3 * c for c in range(Cn)]
    got = L.loglik_batch(pfs, u, y, epochs=epochs)
    assert got.shape == (Cn,) and np.all(np.isfinite(got)) and len(set(got)) == Cn
    for c in (0, 1, 17, Cn - 1):
        assert L.loglik(pfs[c], u, y, epoch=epochs[c]) == got[c]                 # same handle, launched alone
        ref = specs[c].oracle_filter(N, filter=2 if aux else 0, seed=100 + c).loglik(u, y, epoch=epochs[c])
        assert abs(got[c] - ref["ll"]) <= 1e-6 * abs(ref["ll"])
    # state after the batch is the state after loglik: the step verbs continue from it
    assert L.index(pfs[3]) == 61
    assert L.loglik_batch(pfs, u, y, epochs=epochs).tolist() == got.tolist()      # deterministic
    with pytest.raises(L.LLPFError):
        L.loglik_batch([s.particle_filter(N, seed=1)], u, y)                      # not single_block


@pytest.mark.gpu
def test_metropolis_batched_matches_the_posterior(gpu):
    s, u, y = _pmmh_problem(T=200, dims=(2, 2, 2))

    def ffp(theta, pf=None):
        d1, d2 = L.MvNormal(math.exp(theta[0]) * np.eye(2)), L.MvNormal(math.exp(theta[1]) * np.eye(2))
        if pf is None:
            return L.ParticleFilter(1000, L.LinearDynamics(s.A, s.B), L.LinearMeasurement(s.C), d1, d2,
                                    L.MvNormal(s.mu0, s.Sigma0), seed=4, single_block=True)
        return L.set_model(pf, dynamics_density=d1, measurement_density=d2)

    priors = [L.Normal(0, 1.0), L.Normal(0, 1.0)]

    def draw(rng):
        return lambda th: th + 0.1 * rng.standard_normal(2)
    draw.wants_rng = True
    out = L.metropolis_batched(60, ffp, priors, u, y, 160, np.array([0.3, -0.3]), draw, nchains=24, seed=1)
    assert out.shape == (24 * 100, 3) and np.all(np.isfinite(out))
    assert np.all(np.abs(out[:, :2].mean(axis=0)) < 0.6)
    chains = out[:, 0].reshape(24, 100)
    assert np.std(chains[:, -1]) > 0           # the chains are independent (own seeds, own proposals)


@pytest.mark.gpu
@pytest.mark.parametrize("dims,N,T", [((2, 1, 1), 1000, 40), ((4, 2, 2), 5000, 25), ((3, 2, 2), 257, 30)])
def test_device_trajectory_statistics_match_the_host_formulas(gpu, dims, N, T):
    """mean / mode / weighted_cov / weighted_quantile of the solution computed on the device from the HBM-resident history
    (llpf_run_stats) against the same statistics computed on the host from the D2H'd history with the restated StatsBase
    formulas (estimation.py) — and the history itself is in parity with the oracle (test_golden)."""
    s, u, y = _pmmh_problem(T=T, dims=dims)
    pf = s.particle_filter(N, seed=9, resample_threshold=0.5)
    qs = (0.0, 0.05, 0.5, 0.9, 1.0)
    st = L.trajectory_statistics(pf, u, y, q=qs, epoch=3)
    sol = L.forward_trajectory(pf, u, y, epoch=3)
    assert st["ll"] == sol.ll
    assert np.allclose(st["mean"], L.mean_trajectory(sol), rtol=1e-12, atol=1e-13)
    assert np.array_equal(st["mode"], L.mode_trajectory(sol))
    cov = L.weighted_cov(sol)
    for t in range(T):
        assert np.allclose(st["cov"][t], cov[t], rtol=1e-9, atol=1e-12, equal_nan=True), t   # one non-zero weight: 0/0 on both sides
    for k, q in enumerate(qs):
        if q in (0.0, 1.0):
            # At the two ends StatsBase's result is decided by the last bits of its floating-point cumulative sum: at p = 1
            # it compares two differently-ordered sums of the same weights (sum(w) vs the running sum), at p = 0 the walk
            # runs past every particle whose weight is below the double resolution of w_1 (filter weights span hundreds of
            # orders of magnitude).  The device returns the exact answer: the smallest / largest particle value among the
            # non-zero weights.
            for t in range(T):
                nz = sol.we[t] != 0
                ext = sol.x[t][nz].min(axis=0) if q == 0.0 else sol.x[t][nz].max(axis=0)
                assert np.array_equal(st["quantile"][t, k], ext), t
            continue
        ref = L.weighted_quantile(sol, q)
        assert np.allclose(st["quantile"][:, k, :], ref, rtol=1e-9, atol=1e-11), q
    # the per-step weighted mean of the fused reduction agrees with the history-based one
    assert np.allclose(st["mean"], sol.extra["xhat"], rtol=1e-10, atol=1e-12)
