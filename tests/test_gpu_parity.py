"""GPU parity tests: the CUDA path (through the C-ABI / Python host mirror) against the CPU oracle on
identical inputs and identical counter-based RNG streams.

Bars (BASELINE.json north_star): bit-exact particle indices for systematic resampling given the same
u ~ U(0,1); log-likelihood within 1e-6 relative for Float64 (we assert much tighter where the
arithmetic allows).  Full-size cases use size-independent properties (sortedness, exact dyadic
scans, the closed-form Kalman filter).
"""
import numpy as np
import pytest

from models import lg_large_model, lg_model, quadtank_model, ref_model_2state
from oracle import oracle as O

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-6        # the north-star bar
LL_RTOL_TIGHT = 1e-10  # what identical RNG streams + f64 actually give


# ---------------------------------------------------------------------------------------------
# function-boundary parity
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 10, 255, 256, 257, 5000, 100_000])
def test_logsumexp_matches_oracle(gpu, n):
    L = gpu
    w0 = np.random.default_rng(n).standard_normal(n) * 4 - 3
    ll, w, we = L.logsumexp(w0)
    llo, wo, weo = O.logsumexp(w0)
    assert abs(ll - llo) <= 1e-13 * max(1.0, abs(llo))
    assert np.allclose(w, wo, rtol=0, atol=1e-12)
    assert np.allclose(we, weo, rtol=1e-12, atol=1e-300)
    assert abs(we.sum() - 1) < 1e-12


def test_logsumexp_reference_invariants(gpu):   # test/runtests.jl:29-47 on the device path
    L = gpu
    wc = np.random.default_rng(0).standard_normal(10)
    ll, w, we = L.logsumexp(wc)
    assert np.isclose(we.sum(), 1) and np.isclose(np.exp(w).sum(), 1)
    assert np.allclose(w, wc - np.log(np.sum(np.exp(wc))))
    _, wi, _ = L.logsumexp(np.ones(10))
    assert np.allclose(wi, np.full(10, np.log(1 / 10)))


@pytest.mark.parametrize("N,M", [(10, 10), (500, 500), (777, 777), (4096, 4096), (100, 37), (64, 200),
                                 (65536, 65536), (300_001, 300_001)])
def test_systematic_indices_bit_exact_serial_scan(gpu, N, M):
    """Bit-exact indices AND bins against the reference-order oracle (serial cumsum mode)."""
    L = gpu
    rng = np.random.default_rng(N * 7 + M)
    for rep in range(2):
        _, _, we = O.logsumexp(rng.standard_normal(N) * (1 + 2 * rep))
        u = rng.random()
        jo, bo = O.resample_systematic(we, u, M)
        j, b = L.resample(L.ResampleSystematic, we, u, M, scan_mode="serial", return_bins=True)
        assert np.array_equal(b, bo)
        assert np.array_equal(j, jo)


@pytest.mark.parametrize("N", [10, 512, 4096, 65536, 1 << 20])
def test_systematic_indices_bit_exact_fast_scan_dyadic(gpu, N):
    """Dyadic weights: every partial sum is exactly representable, so ANY correct scan must reproduce the
    reference's serial cumsum bit for bit — the fast fixed-point scan does, at full size too."""
    L = gpu
    rng = np.random.default_rng(N)
    k = rng.integers(0, 1 << 20, size=N).astype(np.float64)
    k[rng.integers(0, N, size=max(1, N // 50))] *= 64            # a few heavy particles
    tot = 2.0 ** np.ceil(np.log2(k.sum()))
    we = k / tot                                                  # multiples of 2^-e, sum <= 1
    u = rng.random()
    jo, bo = O.resample_systematic(we, u)
    j, b = L.resample(L.ResampleSystematic, we, u, scan_mode="fast", return_bins=True)
    assert np.array_equal(b, bo)
    assert np.array_equal(j, jo)


@pytest.mark.parametrize("N", [500, 4096, 100_000, 1 << 20])
def test_systematic_fast_scan_flips_are_rounding_ties(gpu, N):
    """Generic weights: the parallel scan re-associates, bins differ from the serial cumsum by O(sqrt(N)) ulp.
    Every index that differs must be a neighbour whose threshold sits inside that rounding gap."""
    L = gpu
    rng = np.random.default_rng(N + 1)
    _, _, we = O.logsumexp(rng.standard_normal(N) * 2)
    u = rng.random()
    jo, bo = O.resample_systematic(we, u)
    j, b = L.resample(L.ResampleSystematic, we, u, scan_mode="fast", return_bins=True)
    assert np.max(np.abs(b - bo)) < 64 * np.sqrt(N) * 2.3e-16
    assert np.all(np.diff(j) >= 0) and j.min() >= 1 and j.max() <= N
    diff = np.nonzero(j != jo)[0]
    assert diff.size <= max(2, N // 20000)
    for i in diff:
        assert abs(int(j[i]) - int(jo[i])) == 1
        s = u * bo[-1] / N + i * (1.0 / N)
        lo = min(j[i], jo[i]) - 1
        assert abs(s - bo[lo]) < 1e-12


def test_systematic_reference_known_answers(gpu):   # test/runtests.jl:88-106 on the device path
    L = gpu
    rng = np.random.default_rng(0)
    _, _, we = L.logsumexp(np.full(10, -np.log(10)))
    for mode in ("fast", "serial"):
        for _ in range(5):
            assert np.array_equal(L.resample(L.ResampleSystematic, we, rng.random(), scan_mode=mode), np.arange(1, 11))
    _, _, we = L.logsumexp(np.array([1., 1, 1, 2, 2, 2, 3, 3, 3]))
    j = L.resample(L.ResampleSystematic, we, rng.random())
    assert j.sum() >= 56 and len(j) == 9
    # stale entries keep the caller's value (resample.jl:26-34)
    j = L.resample(L.ResampleSystematic, np.array([0.25, 0.25, 0.25, 0.2]), 0.9, j0=[9, 9, 9, 9], scan_mode="serial")
    assert list(j) == [1, 2, 3, 9]


@pytest.mark.parametrize("N,M", [(5, 5), (1000, 1000), (5000, 1234), (200_000, 200_000)])
def test_stratified_matches_oracle(gpu, N, M):
    L = gpu
    rng = np.random.default_rng(N + 3 * M)
    _, _, we = O.logsumexp(rng.standard_normal(N))
    u = rng.random(M)
    jo, bo = O.resample_stratified(we, u, M)
    j, b = L.resample(L.ResampleStratified, we, u, M, scan_mode="serial", return_bins=True)
    assert np.array_equal(b, bo) and np.array_equal(j, jo)
    we5 = np.array([0.1, 0.5, 0.1, 0.15, 0.15])            # test/runtests.jl:145-154
    for _ in range(20):
        j = L.resample(L.ResampleStratified, we5, rng.random(5))
        assert j[1] == 2 and j[2] == 2


@pytest.mark.parametrize("N,M", [(5, 5), (10, 10), (777, 777), (4096, 4096), (3000, 1234), (300, 1000), (20_000, 20_000)])
def test_residual_bit_exact_serial_sums(gpu, N, M):
    """resample(ResampleResidual) resample.jl:63-117: bit-exact indices AND bins against the reference-order oracle
    (LLPF_SCAN_SERIAL performs the three sums left to right), draws supplied in the reference's draw order."""
    L = gpu
    rng = np.random.default_rng(11 * N + M)
    for rep in range(2):
        _, _, we = O.logsumexp(rng.standard_normal(N) * (1 + 2 * rep))
        u = rng.random(M)
        j0 = np.full(M, -7, dtype=np.int64)
        jo, bo = O.resample_residual(we, u, M, j0=j0, return_bins=True)
        j, b = L.resample(L.ResampleResidual, we, u, M, j0=j0, scan_mode="serial", return_bins=True)
        assert np.array_equal(b, bo)
        assert np.array_equal(j, jo)
        assert j.min() >= 1 and j.max() <= N          # every slot assigned (no stale entry for generic weights)


@pytest.mark.parametrize("N", [16, 4096, 1 << 16, 1 << 20])
def test_residual_fast_scan_dyadic_and_properties(gpu, N):
    """FAST mode (fixed-point grid sums): with dyadic weights that sum to exactly 1 the counts floor(we*N) and the
    residuals are exact, so the deterministic part must equal the oracle's bit for bit at any size; the multinomial
    part is checked by definition (first i with u < bins[i]) against the device's own bins, at full size too."""
    L = gpu
    rng = np.random.default_rng(N + 5)
    k = rng.integers(0, 1 << 12, size=N).astype(np.float64)
    k[rng.integers(0, N, size=max(1, N // 50))] *= 64
    k[0] += 2.0 ** np.ceil(np.log2(k.sum())) - k.sum()            # sum is a power of two
    we = k / k.sum()
    u = rng.random(N)
    j, b = L.resample(L.ResampleResidual, we, u, scan_mode="fast", return_bins=True)
    cnt = np.floor(we * N).astype(np.int64)
    num = int(cnt.sum())
    assert np.array_equal(j[:num], np.repeat(np.arange(1, N + 1), cnt))
    assert np.all(np.diff(b) >= 0) and b[-1] == 1.0
    resid = we * N - cnt
    assert np.max(np.abs(b - np.cumsum(resid) / resid.sum())) < 1e-12
    assert np.array_equal(j[num:], np.searchsorted(b, u[:N - num], side="right") + 1)
    if N <= 4096:
        jo, bo = O.resample_residual(we, u, N, return_bins=True)
        assert np.max(np.abs(b - bo)) < 1e-13
        assert np.mean(j != jo) < 0.01


def test_residual_no_draw_needed_and_stale_slots(gpu):
    L = gpu
    # every nw is an integer: num == M, early return (resample.jl:85-87), bins holds the (zero) residuals
    we = np.array([0.25, 0.5, 0.0, 0.25])
    for mode in ("serial", "fast"):
        j, b = L.resample(L.ResampleResidual, we, np.full(4, 0.5), scan_mode=mode, return_bins=True)
        assert list(j) == [1, 2, 2, 4] and np.all(b == 0)
    # u >= bins[end] finds no bin: the slot keeps the caller's value (the `break` at :110 never runs)
    we = np.array([0.3, 0.3, 0.4])
    u = np.array([2.0, 0.1, 0.5])
    jo = O.resample_residual(we, u, 3, j0=[9, 9, 9])
    j = L.resample(L.ResampleResidual, we, u, j0=[9, 9, 9], scan_mode="serial")
    assert np.array_equal(j, jo) and 9 in list(j)


@pytest.mark.parametrize("kind", ["systematic", "stratified", "residual"])
def test_resample_proportions_on_device(gpu, kind):     # test/runtests.jl:108-143 (fewer draws: launch cost)
    L = gpu
    we = np.array([0.1, 0.5, 0.1, 0.15, 0.15])
    rng = np.random.default_rng(4)
    counts = np.zeros(5)
    for _ in range(600):
        if kind == "systematic":
            j = L.resample(L.ResampleSystematic, we, rng.random())
        elif kind == "residual":
            j = L.resample(L.ResampleResidual, we, rng.random(5))
        else:
            j = L.resample(L.ResampleStratified, we, rng.random(5))
        counts += np.bincount(j - 1, minlength=5)
    assert np.allclose(counts / counts.sum(), we, atol=0.03)


def test_device_normals_match_oracle_libm(gpu):
    """The device's specialised log/sqrt/sincospi (llpf_math.cuh) against the oracle's libm Box-Muller on the
    same Philox words: reset!(pf) with initial_density = N(0, I) exposes the raw N(0,1) variates."""
    L = gpu
    N = 1 << 18
    for nx in (4, 6):
        pf = L.ParticleFilter(N, L.LinearDynamics(np.eye(nx), np.zeros((nx, 1))), L.LinearMeasurement(np.eye(nx)[:2]),
                              L.MvNormal(np.eye(nx)), L.MvNormal(np.eye(2)), L.MvNormal(np.zeros(nx), np.eye(nx)), seed=99)
        L.reset(pf, 7)
        z = L.particles(pf)
        idx = np.concatenate([np.arange(0, 3000), np.random.default_rng(0).integers(0, N, 3000)])
        zo = np.array([O.normals(99, 7, 0, 0, int(i), nx) for i in idx])
        err = np.abs(z[idx] - zo)
        assert err.max() < 4e-15 * max(1.0, np.abs(zo).max()), err.max()
        assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3 and abs(np.mean(z ** 4) - 3) < 0.05
        assert np.abs(z).max() < 6.8                       # 32-bit uniforms: |z| <= sqrt(-2 ln 2^-33) = 6.76


# ---------------------------------------------------------------------------------------------
# step verbs
# ---------------------------------------------------------------------------------------------
def _assert_state_close(L, pf, of, xtol=1e-11, wtol=1e-10):
    x, xo = L.particles(pf), of.particles
    assert np.allclose(x, xo, rtol=0, atol=xtol), np.abs(x - xo).max()
    assert np.allclose(L.weights(pf), of.weights, rtol=0, atol=wtol)
    assert L.index(pf) == of.index


@pytest.mark.parametrize("nx,nu,ny", [(2, 2, 2), (4, 2, 2), (2, 1, 1), (3, 2, 2), (6, 2, 3)])
def test_reset_and_stepwise_verbs_match_oracle(gpu, nx, nu, ny):
    L = gpu
    s = lg_model(nx, nu, ny, seed=nx * 10 + ny)
    N = 777
    pf = s.particle_filter(N, seed=42, scan_mode="serial", resample_threshold=0.5)
    of = s.oracle_filter(N, seed=42, resample_threshold=0.5)
    L.reset(pf, 3); of.reset(3)
    _assert_state_close(L, pf, of, xtol=1e-13)
    assert np.all(L.weights(pf) == -np.log(N)) and np.all(L.expweights(pf) == 1 / N)
    assert L.num_particles(pf) == N and not L.shouldresample(pf)
    rng = np.random.default_rng(1)
    n_res = 0
    for k in range(25):
        u, y = rng.standard_normal(nu), of.particles.mean(0) @ s.C.T + rng.standard_normal(ny) * 0.5
        t = k * 1.0
        ll, e = L.correct(pf, u, y, None, t)
        llo = of.correct(u, y, t)
        assert e == 0 and abs(ll - llo) <= 1e-11 * max(1, abs(llo))
        assert np.allclose(L.expweights(pf), of.expweights, rtol=1e-9, atol=1e-300)
        assert abs(L.effective_particles(pf) - O.effective_particles(of.expweights)) < 1e-7 * N
        assert L.shouldresample(pf) == of.shouldresample()
        assert np.allclose(L.weighted_mean(pf), of.weighted_mean(), rtol=0, atol=1e-10)
        n_res += of.shouldresample()
        L.predict(pf, u, None, t); of.predict(u, t)
        _assert_state_close(L, pf, of)
        assert np.array_equal(L.ancestors(pf), of.ancestors)
    assert 0 < n_res <= 25


def test_update_and_callable_filter(gpu):
    L = gpu
    s = lg_model(4, 2, 2, seed=1)
    N = 600
    pf = s.particle_filter(N, seed=9, scan_mode="serial")
    of = s.oracle_filter(N, seed=9)
    L.reset(pf, 1); of.reset(1)
    rng = np.random.default_rng(2)
    for k in range(12):
        u, y = rng.standard_normal(2), rng.standard_normal(2)
        # callable filter: t defaults to index(pf)*Ts  (filtering.jl:238)
        ll, _ = pf(u, y)
        llo = of.update(u, y, of.index * 1.0)
        assert abs(ll - llo) <= 1e-11 * max(1, abs(llo))
        _assert_state_close(L, pf, of)


def test_missing_measurement_is_skipped(gpu):   # PFtypes.jl:109
    L = gpu
    s = lg_model()
    pf = s.particle_filter(512, seed=1)
    L.reset(pf, 1)
    L.update(pf, np.zeros(2), np.array([0.3, -0.2]), None, 0.0)
    w = L.weights(pf)
    ll, _ = L.correct(pf, np.zeros(2), np.array([np.nan, 0.0]), None, 1.0)
    assert abs(ll) < 1e-12 and np.allclose(L.weights(pf), w, atol=1e-12)


def test_set_state_roundtrip(gpu):
    L = gpu
    s = lg_model()
    N = 300
    pf = s.particle_filter(N, seed=1, scan_mode="serial")
    of = s.oracle_filter(N, seed=1)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N, 4))
    _, w, _ = O.logsumexp(rng.standard_normal(N) * 3)
    L.set_state(pf, x, w, 5); of.set_state(x, w, 5)
    assert np.array_equal(L.particles(pf), x) and np.allclose(L.weights(pf), w, atol=1e-15)
    u = rng.standard_normal(2)
    assert L.shouldresample(pf) == of.shouldresample()
    L.predict(pf, u, None, 5.0); of.predict(u, 5.0)
    _assert_state_close(L, pf, of)


# ---------------------------------------------------------------------------------------------
# trajectory drivers
# ---------------------------------------------------------------------------------------------
def _data(spec, T, seed, N=64):
    u = np.random.default_rng(seed).standard_normal((T, spec.nu))
    gen = spec.oracle_filter(N, seed=1)
    _, y = gen.simulate(u, seed + 100)
    return u, y


def test_config1_forward_trajectory_full_history(gpu):
    """BASELINE config 1: example_lineargaussian.jl — ParticleFilter, nx=2, N=500, T=200, full x/w/we history."""
    L = gpu
    s = lg_model(2, 2, 2, seed=0)
    N, T = 500, 200
    u, y = _data(s, T, 0)
    pf = s.particle_filter(N, seed=5, scan_mode="serial")
    of = s.oracle_filter(N, seed=5)
    sol = L.forward_trajectory(pf, u, y, epoch=1)
    ref = of.forward_trajectory(u, y, epoch=1, history=True)
    assert sol.x.shape == (T, N, 2) and sol.w.shape == (T, N) and sol.we.shape == (T, N)
    assert np.array_equal(sol.t, np.arange(T) * 1.0)
    assert abs(sol.ll - ref["ll"]) <= LL_RTOL_TIGHT * abs(ref["ll"])
    assert np.array_equal(sol.extra["resampled"], ref["resampled"])
    assert 0 < ref["resampled"].sum() < T
    assert np.allclose(sol.extra["ll_steps"], ref["ll_steps"], rtol=0, atol=1e-10)
    assert np.allclose(sol.extra["ess"], ref["ess"], rtol=1e-9)
    assert np.allclose(sol.x, ref["x"], rtol=0, atol=1e-10)
    assert np.allclose(sol.w, ref["w"], rtol=0, atol=1e-9)
    assert np.allclose(sol.we, ref["we"], rtol=1e-8, atol=1e-300)
    assert np.allclose(sol.extra["xhat"], ref["xhat"], rtol=0, atol=1e-10)
    assert np.allclose(L.mean_trajectory(sol), ref["xhat"], rtol=0, atol=1e-10)
    assert np.allclose(np.sum(sol.we, axis=1), 1.0, atol=1e-12)
    # final state == the oracle's after predict!(T)
    _assert_state_close(L, pf, of)
    # 4-state variant of the same recipe (BASELINE.json calls config 1 "4-state")
    s4 = lg_model(4, 2, 2, seed=0)
    u, y = _data(s4, T, 1)
    pf4, of4 = s4.particle_filter(N, seed=5, scan_mode="serial"), s4.oracle_filter(N, seed=5)
    sol4, ref4 = L.forward_trajectory(pf4, u, y, epoch=2), of4.forward_trajectory(u, y, epoch=2, history=True)
    assert abs(sol4.ll - ref4["ll"]) <= LL_RTOL_TIGHT * abs(ref4["ll"])
    assert np.allclose(sol4.x, ref4["x"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("scan_mode", ["serial", "fast"])
@pytest.mark.parametrize("N", [1000, 16384, 100_000])
def test_loglik_matches_oracle(gpu, N, scan_mode):
    """config-2 model at oracle-affordable sizes; loglik semantics (t = index*Ts)."""
    L = gpu
    s = lg_model(4, 2, 2, seed=0)
    T = 60 if N > 50_000 else 150
    u, y = _data(s, T, 2)
    pf = s.particle_filter(N, seed=77, scan_mode=scan_mode)
    of = s.oracle_filter(N, seed=77)
    got = L.loglik(pf, u, y, epoch=4, details=True)
    ref = of.loglik(u, y, epoch=4)
    assert abs(got["ll"] - ref["ll"]) <= (LL_RTOL_TIGHT if scan_mode == "serial" else LL_RTOL) * abs(ref["ll"])
    assert np.array_equal(got["resampled"], ref["resampled"])
    assert ref["resampled"].sum() > 3
    if scan_mode == "serial":
        _assert_state_close(L, pf, of)


def test_loglik_equals_forward_trajectory_for_time_invariant_model(gpu):
    L = gpu
    s = lg_model()
    u, y = _data(s, 50, 3)
    pf = s.particle_filter(5000, seed=3)
    a = L.loglik(pf, u, y, epoch=9)
    b = L.forward_trajectory(pf, u, y, history=False, epoch=9).ll
    assert a == b
    assert L.loglik(pf, u, y, epoch=9) == a                      # deterministic
    assert L.loglik(pf, u, y, epoch=10) != a                     # a new epoch is a new RNG stream


def test_stratified_in_loop(gpu):
    L = gpu
    s = lg_model()
    N, T = 2000, 80
    u, y = _data(s, T, 5)
    pf = s.particle_filter(N, seed=8, scan_mode="serial", resampling_strategy=L.ResampleStratified)
    of = s.oracle_filter(N, seed=8, resampling=1)
    got, ref = L.loglik(pf, u, y, epoch=1, details=True), of.loglik(u, y, epoch=1)
    assert abs(got["ll"] - ref["ll"]) <= LL_RTOL_TIGHT * abs(ref["ll"])
    assert np.array_equal(got["resampled"], ref["resampled"])
    _assert_state_close(L, pf, of)


@pytest.mark.parametrize("kind", ["pf", "apf", "advanced"])
@pytest.mark.parametrize("scan_mode", ["serial", "fast"])
def test_residual_in_loop(gpu, kind, scan_mode):
    """ResampleResidual inside predict! (filtering.jl:143 / :205): trajectory parity with the oracle; in serial mode
    the ancestors of the last resample are identical."""
    L = gpu
    N, T = 3000, 60
    if kind == "advanced":
        s = quadtank_model()
        u, y = s.inputs(T), None
        of = s.oracle_filter(N, seed=8, resampling=2)
        _, y = of.simulate(u, sim_seed=2)
        pf = s.advanced_filter(N, seed=8, scan_mode=scan_mode, resampling_strategy=L.ResampleResidual)
    else:
        s = lg_model()
        u, y = _data(s, T, 5)
        filt = 2 if kind == "apf" else 0
        of = s.oracle_filter(N, filter=filt, seed=8, resampling=2)
        pf = s.particle_filter(N, seed=8, scan_mode=scan_mode, resampling_strategy=L.ResampleResidual)
        if kind == "apf":
            pf = L.AuxiliaryParticleFilter(pf)
    got, ref = L.loglik(pf, u, y, epoch=1, details=True), of.loglik(u, y, epoch=1)
    assert ref["resampled"].sum() >= 5
    if scan_mode == "serial":
        assert abs(got["ll"] - ref["ll"]) <= LL_RTOL_TIGHT * abs(ref["ll"])
        assert np.array_equal(got["resampled"], ref["resampled"])
        assert np.array_equal(L.ancestors(pf), of.ancestors)
        _assert_state_close(L, pf, of)
    else:
        # re-associated sums may move a count across an integer boundary: a different but equally valid draw
        assert abs(got["ll"] - ref["ll"]) <= 0.02 * abs(ref["ll"])
        j = L.ancestors(pf)
        assert j.min() >= 1 and j.max() <= N


def test_always_resample_threshold_one(gpu):   # resample.jl:6
    L = gpu
    s = lg_model()
    N, T = 1500, 40
    u, y = _data(s, T, 6)
    pf = s.particle_filter(N, seed=8, scan_mode="serial", resample_threshold=1.0)
    of = s.oracle_filter(N, seed=8, resample_threshold=1.0)
    got, ref = L.loglik(pf, u, y, epoch=1, details=True), of.loglik(u, y, epoch=1)
    assert np.all(got["resampled"] == 1) and np.all(ref["resampled"] == 1)
    assert abs(got["ll"] - ref["ll"]) <= LL_RTOL_TIGHT * abs(ref["ll"])
    _assert_state_close(L, pf, of)


def test_advanced_filter_quadtank(gpu):
    """BASELINE config 3 at oracle size: AdvancedParticleFilter, quadtank RK4 (supersample 2), threshold 0.5,
    both time conventions (the t>t_switch leak makes them differ: SURVEY §3.2)."""
    L = gpu
    q = quadtank_model(t_switch=20.0, a1_factor=2.0)
    N, T = 2048, 60
    u = q.inputs(T)
    of = q.oracle_filter(N, seed=4)
    _, y = of.simulate(u, 9)
    pf = q.advanced_filter(N, seed=4, scan_mode="serial")
    assert pf.resample_threshold == 0.5
    sol = L.forward_trajectory(pf, u, y, history=False, epoch=2)
    ref = of.forward_trajectory(u, y, epoch=2)
    assert abs(sol.ll - ref["ll"]) <= 1e-8 * abs(ref["ll"])
    assert np.array_equal(sol.extra["resampled"], ref["resampled"]) and ref["resampled"].sum() > T // 2
    assert np.allclose(sol.extra["xhat"], ref["xhat"], rtol=0, atol=1e-8)
    got, refl = L.loglik(pf, u, y, epoch=2, details=True), of.loglik(u, y, epoch=2)
    assert abs(got["ll"] - refl["ll"]) <= 1e-8 * abs(refl["ll"])
    assert got["ll"] != sol.ll


def test_advanced_filter_lg_equals_particle_filter(gpu):
    L = gpu
    s = lg_model()
    u, y = _data(s, 40, 7)
    a = s.advanced_filter(3000, seed=6, resample_threshold=0.1)
    b = s.particle_filter(3000, seed=6)
    assert L.loglik(a, u, y, epoch=1) == L.loglik(b, u, y, epoch=1)


@pytest.mark.parametrize("N", [500, 20_000])
def test_auxiliary_filter_matches_oracle(gpu, N):
    """BASELINE config 4 model at oracle size: forward_trajectory(pfa) and loglik(pfa) incl. their quirks."""
    L = gpu
    s = lg_model(4, 2, 2, seed=0)
    T = 80
    u, y = _data(s, T, 8)
    pfa = s.aux_filter(N, seed=21, scan_mode="serial")
    ofa = s.oracle_filter(N, filter=2, seed=21)
    hist = N <= 1000
    sol = L.forward_trajectory(pfa, u, y, history=hist, epoch=3)
    ref = ofa.forward_trajectory(u, y, epoch=3, history=hist)
    assert abs(sol.ll - ref["ll"]) <= LL_RTOL_TIGHT * abs(ref["ll"])
    assert abs(sol.extra["ll_steps"][0]) < 1e-12
    assert np.allclose(sol.extra["ll_steps"], ref["ll_steps"], rtol=0, atol=1e-9)
    assert np.array_equal(sol.extra["resampled"], ref["resampled"])
    assert np.allclose(sol.extra["xhat"], ref["xhat"], rtol=0, atol=1e-9)
    if hist:
        assert np.allclose(sol.x, ref["x"], rtol=0, atol=1e-10)
        assert np.allclose(sol.w, ref["w"], rtol=0, atol=1e-9)
        assert np.allclose(sol.we, ref["we"], rtol=1e-8, atol=1e-300)
    got, refl = L.loglik(pfa, u, y, epoch=3, details=True), ofa.loglik(u, y, epoch=3)
    assert abs(got["ll"] - refl["ll"]) <= LL_RTOL_TIGHT * abs(refl["ll"])
    assert np.allclose(got["ll_steps"], refl["ll_steps"], rtol=0, atol=1e-9)
    _assert_state_close(L, pfa, ofa)


def test_auxiliary_stepwise_verbs(gpu):
    L = gpu
    s = lg_model(2, 1, 1, seed=3)
    N = 400
    pfa = s.aux_filter(N, seed=2, scan_mode="serial")
    ofa = s.oracle_filter(N, filter=2, seed=2)
    L.reset(pfa, 1); ofa.reset(1)
    rng = np.random.default_rng(0)
    y = rng.standard_normal((12, 1))
    for k in range(11):
        u = rng.standard_normal(1)
        ll, _ = pfa(u, y[k], y[k + 1], None, k * 1.0)       # update!(pfa,u,y,y1,p,t)
        llo = ofa.update(u, y[k], k * 1.0, y1=y[k + 1])
        assert abs(ll - llo) <= 1e-10 * max(1, abs(llo))
        assert np.allclose(L.particles(pfa), ofa.particles, rtol=0, atol=1e-11)
        assert np.array_equal(L.ancestors(pfa), ofa.ancestors)
    # separate correct! / predict! calls
    ll, _ = L.correct(pfa, np.zeros(1), y[11], None, 11.0)
    assert abs(ll - ofa.correct(np.zeros(1), y[11], 11.0)) < 1e-10
    assert np.allclose(L.weights(pfa), ofa.weights, rtol=0, atol=1e-10)
    L.predict(pfa, np.ones(1), None, 11.0, y1=y[3]); ofa.predict_aux(np.ones(1), y[3], 11.0)
    assert np.allclose(L.particles(pfa), ofa.particles, rtol=0, atol=1e-11)


def test_auxiliary_over_advanced_filter(gpu):   # filtering.jl:219-234: every ll increment is ~0 (Q9)
    L = gpu
    s = lg_model(2, 1, 1, seed=3)
    N, T = 600, 25
    u, y = _data(s, T, 4)
    pfa = L.AuxiliaryParticleFilter(s.advanced_filter(N, seed=2, scan_mode="serial"))
    ofa = s.oracle_filter(N, filter=3, seed=2, resample_threshold=0.5)
    sol = L.forward_trajectory(pfa, u, y, history=False, epoch=1)
    ref = ofa.forward_trajectory(u, y, epoch=1)
    assert np.all(np.abs(sol.extra["ll_steps"]) < 1e-10)
    assert np.allclose(sol.extra["xhat"], ref["xhat"], rtol=0, atol=1e-9)
    assert np.allclose(L.particles(pfa), ofa.particles, rtol=0, atol=1e-10)


# ---------------------------------------------------------------------------------------------
# reference statistical test on the device path + full-size properties
# ---------------------------------------------------------------------------------------------
def test_pf_and_apf_loglik_vs_kalman_on_device(gpu):   # test/runtests.jl:412-449
    L = gpu
    A, B, C_ = ref_model_2state()
    n, T, N = 2, 2000, 1000
    rng = np.random.default_rng(0)
    mu0 = rng.standard_normal(n)
    u = rng.standard_normal((T, 1))

    def omodel(sig):
        return O.ModelArrays(2, 1, 1, C_, sig ** 2 * np.eye(n), np.eye(1), mu0, 4.0 * np.eye(n), A=A, B=B)

    _, y = O.OracleFilter(omodel(0.1), 10, seed=1).simulate(u, 7)
    svec = 10 ** np.linspace(-2, 0, 11)
    llpf, llapf, llkf = [], [], []
    for sg in svec:
        args = (N, L.LinearDynamics(A, B), L.LinearMeasurement(C_), L.MvNormal(np.zeros(n), sg ** 2 * np.eye(n)),
                L.MvNormal(np.eye(1)), L.MvNormal(mu0, 4.0 * np.eye(n)))
        llpf.append(L.loglik(L.ParticleFilter(*args, seed=5), u, y))
        llapf.append(L.loglik(L.AuxiliaryParticleFilter(*args, seed=5), u, y))
        llkf.append(O.kalman_loglik(omodel(sg), u, y))
    llpf, llapf, llkf = map(np.array, (llpf, llapf, llkf))
    assert 4 <= np.argmax(llpf) <= 6 and 4 <= np.argmax(llapf) <= 6 and 4 <= np.argmax(llkf) <= 6
    assert np.max(np.abs(llkf - llpf)) < 20 and np.max(np.abs(llkf - llapf)) < 20


def test_reference_end_to_end_block_on_device(gpu):
    """The assertions of the reference's "End to end" testset (test/runtests.jl:245-333) on the device path: same model,
    N=1000, T=200, M=100; shapes, weighted mean / quantile / covariance consistency, smoothing error bounds for the
    ParticleFilter and the AuxiliaryParticleFilter."""
    L = gpu
    A, B, C_ = ref_model_2state()
    n, N, T, M = 2, 1000, 200, 100
    rng = np.random.default_rng(0)
    mu0 = rng.standard_normal(n)
    args = (N, L.LinearDynamics(A, B), L.LinearMeasurement(C_), L.MvNormal(0.1 ** 2 * np.eye(n)), L.MvNormal(np.eye(1)),
            L.MvNormal(mu0, 4.0 * np.eye(n)))
    pf, pfa = L.ParticleFilter(*args, seed=1), L.AuxiliaryParticleFilter(*args, seed=1)
    assert not L.shouldresample(pf) and not L.shouldresample(pfa)                 # :273-274
    u = rng.standard_normal((T, 1))                                                # du = mvnormal(m, 1)
    om = O.ModelArrays(2, 1, 1, C_, 0.1 ** 2 * np.eye(n), np.eye(1), mu0, 4.0 * np.eye(n), A=A, B=B)
    xs, y = O.OracleFilter(om, 10, seed=1).simulate(u, 7)                          # x,u,y = simulate(pf,T,du)  :279
    sol = L.forward_trajectory(pf, u, y)
    assert sol.x.shape == (T, N, n) and sol.w.shape == (T, N) and sol.we.shape == (T, N) and len(sol.t) == T
    WM = L.mean_trajectory(sol.x, sol.we)                                          # weighted_mean(sol.x, sol.we)  :288
    assert WM.shape == (T, n) and np.array_equal(WM, L.mean_trajectory(sol))       # :289-290
    assert np.allclose(WM[0], (sol.x[0] * sol.we[0][:, None]).sum(axis=0))         # :291
    assert np.allclose(WM, sol.extra["xhat"], rtol=0, atol=1e-9)                   # the fused device reduction agrees
    WQ1, WQ9 = L.weighted_quantile(sol, 0.1), L.weighted_quantile(sol, 0.9)        # :293-297
    assert np.all(WM < WQ9) and np.all(WM > WQ1)
    Cw = L.weighted_cov(sol)                                                       # :299-305
    d = sol.x[1] - WM[1]
    C2 = (d * sol.we[1][:, None]).T @ d
    assert np.allclose(C2 * N / (N - 1), Cw[1])
    xpf, _ = L.mean_trajectory(pf, u, y)                                           # :311
    assert xpf.shape == (T, n)
    xb, ll = L.smooth(pf, M, u, y)                                                 # :313-317
    assert xb.shape == (T, M, n) and np.isfinite(ll)
    xbm = L.smoothed_mean(xb)
    assert np.mean((xs.T - xbm) ** 2) < 5
    xba, _ = L.smooth(pfa, M, u, y)                                                # :319-321
    assert np.mean((xs.T - L.smoothed_mean(xba)) ** 2) < 5
    assert all(np.trace(Cm) < 2 for Cm in L.smoothed_cov(xba))                     # :327-328
    assert L.smoothed_trajs(xba).shape == (n, M, T)                                # :330
    # particle smoothing is at least as good as filtering here (the reference keeps this one as @test_skip, :346)
    assert np.mean((xs - xb.mean(axis=1)) ** 2) < 1.2 * np.mean((xs - WM) ** 2)


# ---------------------------------------------------------------------------------------------
# Float32 particles, wide models (BASELINE config 5; llpf_wide.cuh)
# ---------------------------------------------------------------------------------------------
F32_LL_RTOL = 1e-6     # oracle and device run the same f32 operation sequence; only a f64->f32 double rounding of a
F32_X_ATOL = 1e-5      # normal variate (p ~ 1e-7 per variate) can differ, by one f32 ulp


@pytest.mark.parametrize("nx,nu,ny,offdiag", [(64, 2, 58, 0.0), (64, 2, 58, 0.01), (5, 1, 3, 0.0), (33, 0, 17, 0.02)])
def test_f32_wide_stepwise_verbs_match_oracle(gpu, nx, nu, ny, offdiag):
    L = gpu
    s = lg_large_model(nx, nu, ny, seed=nx + ny, r1_offdiag=offdiag)
    N = 600
    pf = s.particle_filter(N, seed=9, scan_mode="serial", resample_threshold=0.5)
    of = s.oracle_filter(N, seed=9, resample_threshold=0.5)
    L.reset(pf, 2); of.reset(2)
    x0, x0o = L.particles(pf), of.particles
    assert x0.shape == (N, nx) and np.abs(x0 - x0o).max() <= F32_X_ATOL
    assert np.array_equal(x0.astype(np.float32).astype(np.float64), x0)      # values are Float32
    rng = np.random.default_rng(3)
    n_res = 0
    for k in range(8):
        u = rng.standard_normal(nu)
        y = of.particles[0] @ s.C.T + 0.3 * rng.standard_normal(ny)
        ll, _ = L.correct(pf, u, y, None, float(k))
        llo = of.correct(u, y, float(k))
        assert abs(ll - llo) <= F32_LL_RTOL * max(1.0, abs(llo)), (k, ll, llo)
        assert np.allclose(L.expweights(pf), of.expweights, rtol=1e-5, atol=1e-12)
        assert L.shouldresample(pf) == of.shouldresample()
        assert np.allclose(L.weighted_mean(pf), of.weighted_mean(), rtol=0, atol=1e-5)
        n_res += of.shouldresample()
        L.predict(pf, u, None, float(k)); of.predict(u, float(k))
        assert np.array_equal(L.ancestors(pf), of.ancestors)
        x, xo = L.particles(pf), of.particles
        assert np.abs(x - xo).max() <= F32_X_ATOL * max(1.0, np.abs(xo).max()), np.abs(x - xo).max()
        assert (x != xo).mean() < 1e-4          # bit-identical except for the rare double-rounding case
        assert L.index(pf) == of.index
    assert n_res > 0


@pytest.mark.parametrize("scan_mode,thr", [("serial", 0.5), ("fast", 0.5), ("fast", 1.0), ("fast", 0.0)])
def test_f32_wide_loglik_matches_oracle(gpu, scan_mode, thr):
    """config-5 model (64 states, 58 outputs, Float32 particles) at an oracle-affordable size."""
    L = gpu
    s = lg_large_model(seed=1)
    N, T = 2048, 12
    u, y = _data(s, T, 5)
    pf = s.particle_filter(N, seed=21, scan_mode=scan_mode, resample_threshold=thr)
    of = s.oracle_filter(N, seed=21, resample_threshold=thr)
    got = L.loglik(pf, u, y, epoch=1, details=True)
    ref = of.loglik(u, y, epoch=1)
    assert abs(got["ll"] - ref["ll"]) <= F32_LL_RTOL * abs(ref["ll"]), (got["ll"], ref["ll"])
    assert np.array_equal(got["resampled"], ref["resampled"])
    assert np.allclose(got["ll_steps"], ref["ll_steps"], rtol=1e-6, atol=1e-6)
    if scan_mode == "serial":
        assert np.abs(L.particles(pf) - of.particles).max() <= F32_X_ATOL * max(1.0, np.abs(of.particles).max())
    # the same call through forward_trajectory semantics (time-invariant model: same numbers)
    sol = L.forward_trajectory(pf, u, y, epoch=1, history=False)
    assert abs(sol.ll - got["ll"]) <= 1e-12 * abs(got["ll"])


def test_f32_wide_forward_trajectory_history_matches_oracle(gpu):
    """x / w / we history and per-step weighted means of a Float32 wide filter (assembled on the step verbs: the fused loop of
    the wide engine records none) against the oracle's f32 mode; the log-likelihood equals the fused launch's."""
    L = gpu
    s = lg_large_model(seed=3)
    N, T = 512, 8
    u, y = _data(s, T, 6)
    pf = s.particle_filter(N, seed=4, scan_mode="serial", resample_threshold=0.5)
    of = s.oracle_filter(N, seed=4, resample_threshold=0.5)
    sol = L.forward_trajectory(pf, u, y, epoch=2)
    ref = of.forward_trajectory(u, y, epoch=2, history=True)
    assert sol.x.shape == (T, N, s.nx) and sol.w.shape == (T, N)
    assert abs(sol.ll - ref["ll"]) <= F32_LL_RTOL * abs(ref["ll"])
    assert np.array_equal(sol.extra["resampled"], ref["resampled"])
    assert np.abs(sol.x - ref["x"]).max() <= F32_X_ATOL * max(1.0, np.abs(ref["x"]).max())
    assert np.allclose(sol.w, ref["w"], rtol=0, atol=1e-4) and np.allclose(sol.we, ref["we"], rtol=1e-4, atol=1e-9)
    assert np.allclose(L.mean_trajectory(sol), np.einsum("tnd,tn->td", ref["x"], ref["we"]), rtol=0, atol=1e-4)
    fused = L.forward_trajectory(pf, u, y, epoch=2, history=False)
    assert abs(fused.ll - sol.ll) <= 1e-12 * abs(sol.ll)


def test_f32_wide_unsupported_combinations_fail_loudly(gpu):
    L = gpu
    s = lg_large_model(seed=2)
    with pytest.raises(L.LLPFError):
        s.aux_filter(256, seed=1)                       # APF over Float32 particles: not built
    pf = s.particle_filter(256, seed=1)
    u, y = _data(s, 4, 1)
    with pytest.raises(L.LLPFError):
        L.smooth(pf, 8, u, y)                           # the smoother needs the device-resident history
    s8 = lg_model(4, 2, 2, seed=0)
    with pytest.raises(L.LLPFError):
        lg_large_model(65, 2, 4).particle_filter(64)    # nx > 64


# ---------------------------------------------------------------------------------------------
# edge cases: smallest and ragged sizes, no input, weight collapse, a large filter
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,T", [(1, 5), (2, 1), (3, 7), (31, 3), (33, 9), (257, 4), (1000, 1)])
def test_tiny_and_ragged_sizes_match_oracle(gpu, N, T):
    L = gpu
    s = lg_model(3, 1, 2, seed=1)
    u, y = _data(s, T, 2)
    for kind in ("pf", "apf"):
        filt = 2 if kind == "apf" else 0
        of = s.oracle_filter(N, filter=filt, seed=3, resample_threshold=0.5)
        pf = s.particle_filter(N, seed=3, scan_mode="serial", resample_threshold=0.5)
        if kind == "apf":
            pf = L.AuxiliaryParticleFilter(pf)
        got, ref = L.loglik(pf, u, y, epoch=1, details=True), of.loglik(u, y, epoch=1)
        assert abs(got["ll"] - ref["ll"]) <= LL_RTOL_TIGHT * max(1.0, abs(ref["ll"]))
        assert np.array_equal(got["resampled"], ref["resampled"])
        _assert_state_close(L, pf, of)
        sol, refs = L.forward_trajectory(pf, u, y, epoch=2), of.forward_trajectory(u, y, epoch=2, history=True)
        assert sol.x.shape == (T, N, 3) and np.allclose(sol.x, refs["x"], rtol=0, atol=1e-10)
        assert np.allclose(sol.we, refs["we"], rtol=1e-9, atol=1e-300)


def test_model_without_input(gpu):       # nu = 0: dynamics(x,u,p,t) = A*x
    L = gpu
    s = lg_model(4, 0, 2, seed=5)
    T, N = 30, 3000
    from llpf_b200 import workloads as W
    _, y = W.simulate_lg(s, np.zeros((T, 0)), seed=2)
    of = s.oracle_filter(N, seed=3)
    pf = s.particle_filter(N, seed=3, scan_mode="serial")
    got, ref = L.loglik(pf, None, y, epoch=1, details=True), of.loglik(np.zeros((T, 0)), y, epoch=1)
    assert abs(got["ll"] - ref["ll"]) <= LL_RTOL_TIGHT * abs(ref["ll"])
    _assert_state_close(L, pf, of)


def test_weight_collapse_is_reported(gpu):
    """The reference lets NaN propagate silently when every weight underflows (SURVEY §8b error conventions); the C-ABI
    reports it as LLPF_ERR_NONFINITE instead of returning a NaN log-likelihood."""
    L = gpu
    s = lg_model(2, 1, 1, seed=1, r2=1e-12)
    T, N = 6, 512
    u = np.zeros((T, 1))
    y = np.full((T, 1), 1e200)                                 # the quadratic form overflows: every log-weight is -Inf
    pf = s.particle_filter(N, seed=3)
    with pytest.raises(L.LLPFError) as e:
        L.loglik(pf, u, y)
    assert e.value.code == L._abi.ERR_NONFINITE


def test_large_filter_properties(gpu):
    """N = 2^24 (16x the headline size; 1.7 GB of state): the log-likelihood converges to the Kalman filter's, ancestors are
    sorted and in range, weights are finite."""
    L = gpu
    s = lg_model(4, 2, 2, seed=0)
    N, T = 1 << 24, 12
    u = np.random.default_rng(1).standard_normal((T, 2))
    gen = s.oracle_filter(64, seed=1)
    _, y = gen.simulate(u, 4)
    pf = s.particle_filter(N, seed=9, resample_threshold=1.0)
    r = L.loglik(pf, u, y, epoch=1, details=True)
    kf = O.kalman_loglik(s.oracle_model(), u, y)
    assert abs(r["ll"] - kf) < 0.02
    j = L.ancestors(pf)
    assert j.shape == (N,) and j.min() >= 1 and j.max() <= N and np.all(np.diff(j) >= 0)
    assert np.isfinite(L.weights(pf)).all()


# ---------------------------------------------------------------------------------------------
# particle smoother (FFBS)  smoothing.jl:104-143  
# ---------------------------------------------------------------------------------------------
def _traj_mismatch(xb, ref, tol=0.0):
    """fraction of (t, m) entries that differ (a backward draw that flips changes the rest of that trajectory)"""
    return float(np.mean(np.any(np.abs(xb - ref) > tol, axis=2)))


@pytest.mark.parametrize("nx,nu,ny,N,T,M,strategy", [(4, 2, 2, 500, 40, 32, 0), (2, 1, 1, 5000, 30, 100, 0),
                                                      (3, 2, 2, 777, 25, 777, 1), (6, 2, 3, 1000, 20, 50, 2),
                                                      (1, 1, 1, 64, 50, 64, 0), (8, 2, 4, 2048, 12, 33, 0)])
def test_smoother_backward_pass_matches_oracle_on_identical_history(gpu, nx, nu, ny, N, T, M, strategy):
    """smooth(pf, xf, wf, wef, ll, M, u, y) with the ORACLE's forward history as input: the backward simulation on
    identical inputs and identical rand() draws picks the same particles (bit-equal rows of xf)."""
    L = gpu
    s = lg_model(nx, nu, ny, seed=2)
    u, y = _data(s, T, 3)
    strat = [L.ResampleSystematic, L.ResampleStratified, L.ResampleResidual][strategy]
    of = s.oracle_filter(N, seed=6, resampling=strategy)
    sol = of.forward_trajectory(u, y, epoch=4, history=True)
    ref = of.smooth(M, u, sol["x"], sol["w"], sol["we"], epoch=4)
    pf = s.particle_filter(N, seed=6, scan_mode="serial", resampling_strategy=strat)
    xb, ll = L.smooth(pf, sol["x"], sol["w"], sol["we"], sol["ll"], M, u, y, epoch=4)
    assert xb.shape == (T, M, nx) and ll == sol["ll"]
    # a backward draw can only differ from the oracle's when the drawn quantile lies within rounding distance of a bin edge
    # (the device sums the transition densities in another order): no such tie occurs in these fixed-seed cases — every
    # one of the T x M draws picks the oracle's particle
    assert _traj_mismatch(xb, ref) == 0.0
    assert np.array_equal(xb[-1], ref[-1])                     # the resample at T is bit-exact (serial scan)
    assert L.last_smooth_ms(pf) > 0


def test_smoother_end_to_end_matches_oracle(gpu):
    """xb, ll = smooth(pf, M, u, y): forward pass on the device (history stays in HBM) + backward simulation."""
    L = gpu
    s = lg_model(4, 2, 2, seed=0)
    N, T, M = 2000, 60, 100                                    # test/runtests.jl:264-333 uses N=2000, M=100
    u, y = _data(s, T, 9)
    of = s.oracle_filter(N, seed=3)
    sol = of.forward_trajectory(u, y, epoch=2, history=True)
    ref = of.smooth(M, u, sol["x"], sol["w"], sol["we"], epoch=2)
    pf = s.particle_filter(N, seed=3, scan_mode="serial")
    xb, ll = L.smooth(pf, M, u, y, epoch=2)
    assert abs(ll - sol["ll"]) <= LL_RTOL_TIGHT * abs(sol["ll"])
    assert _traj_mismatch(xb, ref, tol=1e-9) == 0.0
    # helpers smoothing.jl:350-385
    assert L.smoothed_mean(xb).shape == (4, T) and L.smoothed_trajs(xb).shape == (4, M, T)
    assert len(L.smoothed_cov(xb)) == T and L.smoothed_cov(xb)[0].shape == (4, 4)
    # APF forward pass + the same backward simulation (test/runtests.jl:319)
    apf = L.AuxiliaryParticleFilter(s.particle_filter(N, seed=3, scan_mode="serial"))
    oa = s.oracle_filter(N, filter=2, seed=3)
    sola = oa.forward_trajectory(u, y, epoch=2, history=True)
    refa = oa.smooth(M, u, sola["x"], sola["w"], sola["we"], epoch=2)
    xba, lla = L.smooth(apf, M, u, y, epoch=2)
    assert abs(lla - sola["ll"]) <= LL_RTOL_TIGHT * abs(sola["ll"])
    assert _traj_mismatch(xba, refa, tol=1e-9) == 0.0


def test_smoother_quadtank_and_errors(gpu):
    L = gpu
    q = quadtank_model()
    N, T, M = 1024, 30, 40
    u = q.inputs(T)
    of = q.oracle_filter(N, seed=4)
    _, y = of.simulate(u, 9)
    sol = of.forward_trajectory(u, y, epoch=2, history=True)
    ref = of.smooth(M, u, sol["x"], sol["w"], sol["we"], epoch=2)
    pf = q.advanced_filter(N, seed=4, scan_mode="serial")
    xb, _ = L.smooth(pf, sol["x"], sol["w"], sol["we"], sol["ll"], M, u, y, epoch=2)
    assert _traj_mismatch(xb, ref) == 0.0
    with pytest.raises(L.LLPFError):                           # @assert M <= N   smoothing.jl:122
        L.smooth(pf, N + 1, u, y)


def test_smoother_large_properties(gpu):
    """Beyond oracle size (N = 2^16, M = 256): every smoothed state is a filtered particle of its step, trajectories are
    distinct draws, and the smoothed mean is at least as close to the truth as the filter mean."""
    L = gpu
    s = lg_model(4, 2, 2, seed=0)
    N, T, M = 1 << 16, 40, 256
    u = np.random.default_rng(2).standard_normal((T, 2))
    gen = s.oracle_filter(64, seed=1)
    xs, y = gen.simulate(u, 11)
    pf = s.particle_filter(N, seed=3)
    sol = L.forward_trajectory(pf, u, y, epoch=5)
    xb, ll = L.smooth(pf, sol.x, sol.w, sol.we, sol.ll, M, u, y, epoch=5)
    for t in (0, T // 2, T - 1):
        keys = {row.tobytes() for row in sol.x[t]}
        assert all(xb[t, m].tobytes() in keys for m in range(M))
    assert len({xb[0, m].tobytes() for m in range(M)}) > M // 4
    err_f = np.mean((L.mean_trajectory(sol) - xs) ** 2)
    err_s = np.mean((xb.mean(axis=1) - xs) ** 2)
    assert err_s < 1.05 * err_f
    xb2, ll2 = L.smooth(pf, M, u, y, epoch=5)                  # same seed/epoch -> same forward pass -> same draws
    assert ll2 == sol.ll and np.array_equal(xb2, xb)


def test_full_size_config2_properties(gpu):
    """N = 2^20 (BASELINE config 2 particle count), T = 100: the oracle is too slow here, so check
    size-independent properties: the log-likelihood converges to the closed-form Kalman filter
    (std ~ sqrt(c*T/N)), ancestors are sorted and in range, weights normalise, runs are deterministic."""
    L = gpu
    s = lg_model(4, 2, 2, seed=0)
    N, T = 1 << 20, 100
    u, y = _data(s, T, 11)
    pf = s.particle_filter(N, seed=123)
    r = L.loglik(pf, u, y, epoch=1, details=True)
    kf = O.kalman_loglik(s.oracle_model(), u, y)
    assert abs(r["ll"] - kf) < 0.15, (r["ll"], kf)
    assert 0 < r["resampled"].sum() < T
    assert np.all(r["ess"] > 1) and np.all(r["ess"] <= N * (1 + 1e-9))
    j = L.ancestors(pf)
    assert j.min() >= 1 and j.max() <= N and (np.all(np.diff(j) >= 0) or r["resampled"][-1] == 0)
    assert abs(L.expweights(pf).sum() - 1) < 1e-10
    assert L.loglik(pf, u, y, epoch=1) == r["ll"]
    x = L.particles(pf)
    assert np.all(np.isfinite(x))


# ---------------------------------------------------------------------------------------------
# round 2: headline-size oracle parity, Julia's range paths, multi-GPU inside pytest
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scan_mode", ["serial", "fast"])
def test_headline_size_config2_against_oracle(gpu, scan_mode):
    """BASELINE config 2 at its full particle count (N = 2^20, the bench.py workload) for the first T = 40 time steps,
    DIRECTLY against the CPU oracle (10 s of oracle time).
    SERIAL scan (the reference's summation order): log-likelihood to 1e-10, every resample decision, per-step ll / ESS,
    ancestors bit-exact, particles to 1e-10.
    FAST scan (what bench.py times): the fixed-point scan rounds `bins` differently from the serial f64 cumsum by
    O(sqrt(N)) ulp, so about N^2 * 1e-14 ~ 0.01-1 thresholds per resample fall on the other side of a bin edge.  Systematic
    resampling is chaotic in that respect: ONE moved ancestor shifts every later bin edge of the next resample by ~1e-6 = one
    threshold spacing, after which the two runs are different (statistically equivalent) realisations.  So FAST is held to:
    identical to the oracle (1e-10) up to the first moved ancestor, never diverging before the first resample, and within
    Monte-Carlo distance of the oracle and of the closed-form Kalman filter afterwards (measured on B200: identical for
    the first 30 of the 40 steps = 13 resamples, then |ll - ll_oracle| = 5.9e-3).  (The per-resample claim — every
    moved ancestor is a +-1 neighbour inside the rounding gap — is test_systematic_fast_scan_flips_are_rounding_ties.)"""
    L = gpu
    from llpf_b200 import workloads as W
    s = lg_model(4, 2, 2, seed=0)
    N, T = 1 << 20, 40
    u = np.random.default_rng(0).standard_normal((T, 2))
    _, y = W.simulate_lg(s, u, seed=1)                 # bench.py's workload(T) prefix
    of = s.oracle_filter(N, filter=0, resample_threshold=0.1, seed=1)
    ref = of.loglik(u, y, epoch=1)
    assert 0 < ref["resampled"].sum() < T
    pf = s.particle_filter(N, seed=1, resample_threshold=0.1, scan_mode=scan_mode)
    got = L.loglik(pf, u, y, epoch=1, details=True)
    scale = max(1.0, abs(ref["ll"]))
    if scan_mode == "serial":
        assert abs(got["ll"] - ref["ll"]) <= LL_RTOL_TIGHT * abs(ref["ll"])
        assert np.array_equal(got["resampled"], ref["resampled"])
        assert np.allclose(got["ll_steps"], ref["ll_steps"], rtol=0, atol=1e-9 * scale)
        assert np.allclose(got["ess"], ref["ess"], rtol=1e-9)
        assert np.array_equal(L.ancestors(pf), of.ancestors)
        assert np.abs(L.particles(pf) - of.particles).max() <= 1e-10
        assert np.abs(L.weights(pf) - of.weights).max() <= 1e-9
        return
    d = np.abs(got["ll_steps"] - ref["ll_steps"])
    moved = np.nonzero(d > 1e-9 * scale)[0]
    first_res = int(np.nonzero(ref["resampled"])[0][0])
    k = int(moved[0]) if moved.size else T
    print(f"FAST scan, N=2^20: identical to the oracle for the first {k} of {T} steps "
          f"({int(ref['resampled'][:k].sum())} resamples); |ll - ll_oracle| = {abs(got['ll'] - ref['ll']):.3e}")
    assert k > first_res                                # nothing can differ before an ancestor has been chosen
    assert np.array_equal(got["resampled"][:k], ref["resampled"][:k])
    assert np.allclose(got["ess"][:k], ref["ess"][:k], rtol=1e-9)
    # afterwards: another realisation of the same estimator (std of ll at N = 2^20, T = 40 is ~ 0.1: measured 0.27 at
    # N = 2^17 over seeds; the two runs share their first k steps, so they stay much closer to each other than to the KF)
    kf = O.kalman_loglik(s.oracle_model(), u, y)
    assert abs(got["ll"] - ref["ll"]) < 0.1 and abs(got["ll"] - kf) < 0.5
    if k == T:
        assert abs(got["ll"] - ref["ll"]) <= LL_RTOL * abs(ref["ll"])


@pytest.mark.parametrize("scan_mode", ["serial", "fast"])
def test_systematic_follows_julia_rational_range(gpu, scan_mode):
    """rand() values for which Julia builds the threshold range r:(1/M):(bins[N]+r) from exact integer ratios (double-double
    elements, one ulp away from fl(r + i/M) for non-dyadic M): the stand-alone entry must select the same indices as the
    oracle, which follows base/twiceprecision.jl (oracle/julia_range.py, tests/test_julia_range.py)."""
    L = gpu
    from oracle import julia_range as J
    hit = 0
    for N in (5, 10, 7, 12, 25, 100, 1000, 777):
        we = np.full(N, 1.0 / N)
        for u in (0.0, 0.5, 0.25):
            jo, bo = O.resample_systematic(we, u)
            j, b = L.resample(L.ResampleSystematic, we, u, scan_mode=scan_mode, return_bins=True)
            if scan_mode == "serial":
                assert np.array_equal(b, bo)
            R, _ = O.julia_range(u * bo[-1] / N, 1.0 / N, bo[-1] + u * bo[-1] / N)
            hit += R.rational
            if scan_mode == "serial" or np.array_equal(b, bo):
                assert np.array_equal(j, jo), (N, u)
    assert hit >= 6
    j5 = L.resample(L.ResampleSystematic, np.full(5, 0.2), 0.0, scan_mode="serial")
    assert list(j5) == [1, 2, 3, 3, 5]       # Julia's rational range; the literal formula would give 1:5
    # dyadic weights: every partial sum exact, so FAST == SERIAL == oracle on both range paths
    rng = np.random.default_rng(3)
    for N, M in ((96, 96), (1000, 1000), (384, 100)):
        k = rng.integers(1, 64, N).astype(np.float64)
        we = k / 2.0 ** 16                      # dyadic, un-normalised (sum < 1)
        for u in (0.0, 0.5, 0.125, float(rng.random())):
            jo, bo = O.resample_systematic(we, u, M)
            j, b = L.resample(L.ResampleSystematic, we, u, M, scan_mode=scan_mode, return_bins=True)
            assert np.array_equal(b, bo) and np.array_equal(j, jo), (N, M, u)


def test_sharded_filters_inside_pytest(gpu):
    """tests/multi_gpu_worker.py (sharded PF / APF / Float32-wide filters: bit-identical to one GPU, oracle parity,
    sharded accessors) run under torchrun on 2 GPUs and on every GPU of the box — part of `pytest -m gpu` so the driver
    verifies the multi-GPU numbers are numbers of a correct filter.  Skips (with the reason printed) on a 1-GPU box."""
    import ctypes as C
    import os
    import subprocess
    import sys
    L = gpu
    n = C.c_int()
    L.load_library().llpf_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip(f"multi-GPU parity needs >= 2 GPUs on the box (found {n.value})")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    worlds = sorted({2, min(8, n.value)})
    for world in worlds:
        port = 29600 + world
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "multi_gpu_worker.py")]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
        print(r.stdout[-4000:])
        assert r.returncode == 0, r.stderr[-3000:]
        assert "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_forward_trajectory_callbacks(gpu):
    """The four callbacks of forward_trajectory (filtering.jl:343,353-362): called in the reference's order with the
    reference's arguments, and the stepwise loop reproduces the fused single-launch trajectory."""
    L = gpu
    s = lg_model(4, 2, 2, seed=0)
    N, T = 2000, 25
    u, y = _data(s, T, 4)
    pf = s.particle_filter(N, seed=6, resample_threshold=0.5)
    fused = L.forward_trajectory(pf, u, y, epoch=2)
    log = []
    sol = L.forward_trajectory(
        pf, u, y, epoch=2,
        pre_correct_cb=lambda f, ut, yt, p, t: log.append(("pre_correct", t)),
        post_correct_cb=lambda f, ut, yt, p, t, ll: log.append(("post_correct", t, ll, L.effective_particles(f))),
        pre_predict_cb=lambda f, ut, yt, p, t, ll: log.append(("pre_predict", t, ll)),
        post_predict_cb=lambda f, ut, yt, p, t: log.append(("post_predict", t, L.index(f))))
    assert [e[0] for e in log[:4]] == ["pre_correct", "post_correct", "pre_predict", "post_predict"] and len(log) == 4 * T
    assert [e[1] for e in log[::4]] == [float(k) for k in range(T)]
    assert log[3][2] == 2 and log[-1][2] == T + 1                         # index(pf) after predict!
    assert abs(sol.ll - fused.ll) <= 1e-12 * abs(fused.ll)
    assert np.array_equal(sol.extra["resampled"], fused.extra["resampled"])
    assert np.allclose(sol.x, fused.x, rtol=0, atol=1e-12) and np.allclose(sol.we, fused.we, rtol=1e-10, atol=1e-300)
    assert np.allclose([e[2] for e in log[1::4]], fused.extra["ll_steps"], rtol=0, atol=1e-10)
    assert np.allclose([e[3] for e in log[1::4]], fused.extra["ess"], rtol=1e-9)
    # a callback that intervenes: freezing the particles before every predict! changes the result
    def freeze(f, ut, yt, p, t, ll):
        L.set_state(f, np.zeros((N, 4)), L.weights(f), L.index(f))
    sol2 = L.forward_trajectory(pf, u, y, epoch=2, pre_predict_cb=freeze)
    assert abs(sol2.ll - fused.ll) > 1e-3


def test_metropolis_resampling_extension(gpu):
    """Metropolis resampling is NOT in the reference (north-star extension): statistical tests only, in the style of
    test/runtests.jl:108-143 — empirical proportions of the drawn indices match the weights, and a ParticleFilter that
    uses it estimates the same log-likelihood as the systematic one / the Kalman filter."""
    L = gpu
    we = np.array([0.1, 0.5, 0.1, 0.15, 0.15])
    j = L.resample(L.ResampleMetropolis, we, 12345, M=20000)
    assert j.min() >= 1 and j.max() <= 5
    prop = np.bincount(j, minlength=6)[1:] / j.size
    assert np.all(np.abs(prop - we) < 0.02), prop
    # one heavy particle: the chains need enough proposals to find it (the known weakness of the method: B must grow with
    # the largest weight ratio) — B = 32 proposals among N = 20 indices do, among N = 1000 they would not
    w2 = np.full(20, 0.5 / 19); w2[7] = 0.5
    j2 = L.resample(L.ResampleMetropolis, w2, 7, M=20000)
    assert abs(np.mean(j2 == 8) - 0.5) < 0.03
    s = lg_model(4, 2, 2, seed=0)
    N, T = 1 << 16, 60
    u, y = _data(s, T, 2)
    kf = O.kalman_loglik(s.oracle_model(), u, y)
    lls = {}
    for name, strat in (("sys", L.ResampleSystematic), ("met", L.ResampleMetropolis)):
        vals = []
        for seed in range(4):
            pf = s.particle_filter(N, seed=seed, resample_threshold=0.5, resampling_strategy=strat, metropolis_steps=48)
            r = L.loglik(pf, u, y, epoch=1, details=True)
            vals.append(r["ll"])
            assert r["resampled"].sum() > 5
        lls[name] = np.array(vals)
    assert abs(lls["met"].mean() - lls["sys"].mean()) < 4 * (lls["sys"].std() + lls["met"].std() + 0.05)
    assert abs(lls["met"].mean() - kf) < 1.5
    anc = L.ancestors(pf)                      # state.j of the last Metropolis resample: valid global indices, not sorted
    assert anc.min() >= 1 and anc.max() <= N
    # APF and history work with it too
    pfa = L.AuxiliaryParticleFilter(s.particle_filter(4096, seed=1, resampling_strategy=L.ResampleMetropolis))
    sol = L.forward_trajectory(pfa, u, y, epoch=1)
    assert np.isfinite(sol.ll) and np.allclose(sol.we.sum(axis=1), 1.0)
