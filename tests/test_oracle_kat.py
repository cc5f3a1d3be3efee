"""Pins the CPU oracle (oracle/llpf_oracle.c) against every known-answer / invariant test the
reference's own suite holds for the particle-filter path, plus independent
re-derivations in numpy of the quirks the restatement must keep (Q1-Q12)."""
import numpy as np
import pytest

from models import lg_model, quadtank_model, ref_model_2state
from oracle import oracle as O


# ---- RNG contract -------------------------------------------------------------------------------
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normals_moments_and_counter_independence():
    z = np.concatenate([O.normals(5, 0, 1, t, i, 4) for t in range(40) for i in range(400)])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    assert abs(np.mean(z ** 4) - 3) < 0.2
    # same counter -> same variate; different epoch / stream / step / particle -> different
    a = O.normals(5, 0, 1, 3, 7, 4)
    assert np.array_equal(a, O.normals(5, 0, 1, 3, 7, 4))
    for other in (O.normals(5, 1, 1, 3, 7, 4), O.normals(5, 0, 0, 3, 7, 4), O.normals(5, 0, 1, 4, 7, 4),
                  O.normals(5, 0, 1, 3, 8, 4), O.normals(6, 0, 1, 3, 7, 4)):
        assert not np.array_equal(a, other)
    # nx > 4 uses a second Philox block, the first four are unchanged
    assert np.array_equal(O.normals(5, 0, 1, 3, 7, 6)[:4], a)
    u = O.uniform53(1, 0, 2, 9, 0)
    assert 0.0 <= u < 1.0 and (u * 2 ** 53) == int(u * 2 ** 53)


# ---- test/runtests.jl:29-47 "logsumexp" ---------------------------------------------------------
def test_logsumexp_invariants():
    rng = np.random.default_rng(0)
    wc = rng.standard_normal(10)
    ll, w, we = O.logsumexp(wc)
    assert np.isclose(we.sum(), 1)
    assert np.isclose(np.exp(w).sum(), 1)
    assert np.allclose(w, wc - np.log(np.sum(np.exp(wc))))
    assert np.isclose(ll, np.log(np.sum(np.exp(wc))))
    _, wi, _ = O.logsumexp(np.ones(10))
    assert np.allclose(wi, np.full(10, np.log(1 / 10)))
    wc = rng.standard_normal(10)
    we, w2 = O.expnormalize(wc, inplace=False)
    assert np.isclose(we.sum(), 1)
    assert np.allclose(w2, wc, atol=1e-15, rtol=0)
    assert np.isclose(O.expnormalize(wc).sum(), 1)


def test_logsumexp_first_argmax_and_large_n():
    # Q5: first arg-max is the one excluded from the sum; ties must not double count
    w = np.array([0.0, 3.0, 3.0, -1.0])
    ll, wn, we = O.logsumexp(w)
    assert np.isclose(ll, np.log(np.exp(w).sum()))
    n = 5000  # exercises the pairwise split (> 1024)
    w = np.random.default_rng(1).standard_normal(n) * 3
    ll, wn, we = O.logsumexp(w)
    assert np.isclose(ll, np.log(np.sum(np.exp(w - w.max()))) + w.max(), rtol=1e-14)
    assert abs(we.sum() - 1) < 1e-13


# ---- test/runtests.jl:88-106 "resample systematic" ----------------------------------------------
def test_resample_systematic_known_answers():
    N = 10
    we = np.full(N, 1 / N)
    assert np.isclose(O.effective_particles(we), 10)
    _, _, we = O.logsumexp(np.full(N, -np.log(N)))
    rng = np.random.default_rng(0)
    for _ in range(20):
        j, _ = O.resample_systematic(we, rng.random())
        assert np.array_equal(j, np.arange(1, 11))
    _, _, we = O.logsumexp(np.array([1., 1, 1, 2, 2, 2, 3, 3, 3]))
    for _ in range(20):
        j, _ = O.resample_systematic(we, rng.random())
        assert j.sum() >= 56 and len(j) == len(we)
    for _ in range(10):
        _, _, we = O.logsumexp(rng.standard_normal(100))
        j, _ = O.resample_systematic(we, rng.random())
        assert j.max() <= 100 and j.min() >= 1
        assert np.all(np.diff(j) >= 0)


# ---- test/runtests.jl:108-143 "resample correct proportions" ------------------------------------
@pytest.mark.parametrize("kind", ["systematic", "stratified", "residual"])
def test_resample_proportions(kind):
    we = np.array([0.1, 0.5, 0.1, 0.15, 0.15])
    K, R = 5, 10000
    rng = np.random.default_rng(2)
    counts = np.zeros(K)
    for _ in range(R):
        if kind == "systematic":
            j, _ = O.resample_systematic(we, rng.random())
        elif kind == "stratified":
            j, _ = O.resample_stratified(we, rng.random(K))
        else:
            j = O.resample_residual(we, rng.random(K))
        counts += np.bincount(j - 1, minlength=K)
    assert np.allclose(counts / counts.sum(), we, atol=0.02)


# ---- test/runtests.jl:145-154 "resample stratified" ---------------------------------------------
def test_resample_stratified_known_answer():
    we = np.array([0.1, 0.5, 0.1, 0.15, 0.15])
    rng = np.random.default_rng(3)
    for _ in range(100):
        j, _ = O.resample_stratified(we, rng.random(5))
        assert j[1] == 2 and j[2] == 2


# ---- independent re-derivation of resample.jl:17-36 in plain Python floats ----------------------
def _py_systematic(we, u01, M=None, j0=None):
    N = len(we)
    M = N if M is None else M
    bins = [0.0] * N
    bins[0] = float(we[0])
    for i in range(1, N):
        bins[i] = bins[i - 1] + float(we[i])          # Q4 serial cumsum
    r = u01 * bins[-1] / N                            # Q1
    step = 1.0 / M
    j = list(range(1, M + 1)) if j0 is None else list(j0)
    bo = 0
    for i in range(M):
        s = r + (i * step)                            # Q2: product rounded, then the sum
        for b in range(bo, N):
            if s < bins[b]:
                j[i] = b + 1
                bo = b
                break                                 # Q3: no hit -> j[i] untouched
    return np.array(j), np.array(bins)


@pytest.mark.parametrize("N,M", [(10, 10), (500, 500), (777, 777), (100, 37), (64, 200)])
def test_resample_systematic_matches_python_restatement(N, M):
    rng = np.random.default_rng(N + M)
    for _ in range(5):
        _, _, we = O.logsumexp(rng.standard_normal(N) * 2)
        u = rng.random()
        j, b = O.resample_systematic(we, u, M)
        jp, bp = _py_systematic(we, u, M)
        assert np.array_equal(b, bp)
        assert np.array_equal(j, jp)


# ---- independent re-derivations of resample.jl:38-61 (stratified) and :63-117 (residual) in plain Python floats ------
def _py_stratified(we, us, M=None, j0=None):
    N = len(we)
    M = N if M is None else M
    bins = [0.0] * N
    bins[0] = float(we[0])
    for i in range(1, N):
        bins[i] = bins[i - 1] + float(we[i])
    j = [0] * M if j0 is None else list(j0)
    bo = 0
    for i in range(M):
        s = ((i + float(us[i])) / M) * bins[-1]        # ((i-1) + rand())/M * bins[end], i 1-based in the reference
        for b in range(bo, N):
            if s < bins[b]:
                j[i] = b + 1
                bo = b
                break
    return np.array(j), np.array(bins)


def _py_residual(we, us, M=None, j0=None):
    N = len(we)
    M = N if M is None else M
    wsum = 0.0
    for i in range(N):
        wsum += float(we[i])
    inv_wsum = 1 / wsum
    j = [0] * M if j0 is None else list(j0)
    bins = [0.0] * N
    num = 0
    for i in range(N):
        nw = float(we[i]) * inv_wsum * M
        cnt = int(np.floor(nw))
        bins[i] = nw - cnt
        for _ in range(cnt):
            j[num] = i + 1
            num += 1
    if num == M:
        return np.array(j), np.array(bins)
    rsum = 0.0
    for i in range(N):
        rsum += bins[i]
    inv_rsum = 1 / rsum
    for i in range(N):
        bins[i] *= inv_rsum
    for i in range(1, N):
        bins[i] += bins[i - 1]
    k = 0
    for m in range(num, M):
        u = float(us[k]); k += 1
        for i in range(N):
            if u < bins[i]:
                j[m] = i + 1
                break
    return np.array(j), np.array(bins)


@pytest.mark.parametrize("N,M", [(5, 5), (10, 10), (300, 300), (100, 37), (64, 200)])
def test_resample_stratified_and_residual_match_python_restatements(N, M):
    rng = np.random.default_rng(3 * N + M)
    for rep in range(4):
        _, _, we = O.logsumexp(rng.standard_normal(N) * (1 + rep))
        us = rng.random(M)
        j0 = np.full(M, -3, dtype=np.int64)
        j, b = O.resample_stratified(we, us, M, j0=j0)
        jp, bp = _py_stratified(we, us, M, j0=j0)
        assert np.array_equal(b, bp) and np.array_equal(j, jp)
        j, b = O.resample_residual(we, us, M, j0=j0, return_bins=True)
        jp, bp = _py_residual(we, us, M, j0=j0)
        assert np.array_equal(b, bp) and np.array_equal(j, jp)
        assert np.all(np.sort(j[j > 0]) >= 1) and j.max() <= N


def test_resample_stale_entries_keep_previous_value():
    # Q3: weights that sum to less than the last threshold leave trailing j untouched
    we = np.array([0.25, 0.25, 0.25, 0.2])       # bins[end] = 0.95
    j0 = np.array([9, 9, 9, 9])
    j, _ = O.resample_systematic(we, 0.9, j0=j0)  # r = 0.9*0.95/4 ; s[4] = r + 0.75 = 0.96375 > 0.95
    assert j[3] == 9 and list(j[:3]) == [1, 2, 3]


# ---- test/runtests.jl:182-188 "rk4" + quadtank ----------------------------------------------------
def test_rk4_known_answers():
    assert np.isclose(O.rk4_constant_rhs(-1.0, 1.0, 1.0), 0.0)
    assert np.isclose(O.rk4_constant_rhs(-1.0, 0.0, 1.0), -1.0)
    assert np.isclose(O.rk4_constant_rhs(-1.0, 1.0, 1.0, supersample=4), 0.0)


def _np_quadtank(h, u, p, t, t_switch=np.inf, f=1.0):
    kc, k1, k2, A, a, g = p
    a1 = a * (f if t > t_switch else 1.0)
    ss = lambda v: np.sqrt(max(v, 0.0) + 1e-3)  # noqa: E731
    tg = 2 * 9.81
    return np.array([
        -a1 / A * ss(tg * h[0]) + a / A * ss(tg * h[2]) + g * k1 / A * u[0],
        -a / A * ss(tg * h[1]) + a / A * ss(tg * h[3]) + g * k2 / A * u[1],
        -a / A * ss(tg * h[2]) + (1 - g) * k2 / A * u[1],
        -a / A * ss(tg * h[3]) + (1 - g) * k1 / A * u[0]])


def _np_rk4(fun, x, u, t, Ts0, supersample):
    Ts = Ts0 / supersample
    for _ in range(supersample):
        f1 = fun(x, u, t)
        f2 = fun(x + Ts / 2 * f1, u, t + Ts / 2)
        f3 = fun(x + Ts / 2 * f2, u, t + Ts / 2)
        f4 = fun(x + Ts * f3, u, t + Ts)
        x = x + Ts / 6 * (f1 + 2 * f2 + 2 * f3 + f4)
        t += Ts
    return x


def test_quadtank_rk4_against_numpy():
    q = quadtank_model(t_switch=500.0, a1_factor=2.0)
    of = q.oracle_filter(4)
    rng = np.random.default_rng(0)
    for t in (0.0, 499.0, 499.6, 500.0, 731.0):   # the t>500 switch flips inside / between sub-steps
        x = np.abs(rng.standard_normal(4)) * 3
        u = rng.random(2)
        got = of.dynamics(x, u, t)
        exp = _np_rk4(lambda h, uu, tt: _np_quadtank(h, uu, q.p, tt, 500.0, 2.0), x, u, t, 1.0, 2)
        assert np.allclose(got, exp, rtol=1e-14, atol=0)


# ---- step semantics -----------------------------------------------------------------------------
def test_reset_and_uniform_weights():
    s = lg_model()
    of = s.oracle_filter(1000, seed=11)
    of.reset(3)
    assert of.index == 1
    assert np.all(of.weights == -np.log(1000)) and np.all(of.expweights == 1 / 1000)
    assert np.array_equal(of.particles, of.xprev)
    x = of.particles
    assert np.allclose(x.mean(0), s.mu0, atol=0.3) and np.allclose(np.cov(x.T), s.Sigma0, atol=0.6)
    assert not of.shouldresample()


def test_correct_predict_manual_equivalence():
    """update! == correct! then predict! (filtering.jl:181-185) and the bootstrap step written out in numpy"""
    s = lg_model(nx=2, nu=2, ny=2, seed=4)
    N = 300
    a, b = s.oracle_filter(N, seed=2), s.oracle_filter(N, seed=2)
    a.reset(1); b.reset(1)
    rng = np.random.default_rng(0)
    for k in range(15):
        u, y = rng.standard_normal(2), rng.standard_normal(2)
        x0, w0 = a.particles.copy(), a.weights.copy()
        lla = a.update(u, y, k * 1.0)
        llb = b.correct(u, y, k * 1.0)
        # numpy restatement of correct!
        r = y[None, :] - x0 @ s.C.T
        w1 = w0 + (-(2 * np.log(2 * np.pi)) / 2 - 0.5 * np.sum(r * r, axis=1))
        assert np.isclose(llb, np.log(np.sum(np.exp(w1 - w1.max()))) + w1.max(), rtol=1e-13)
        assert np.allclose(b.weights, w1 - llb, rtol=0, atol=1e-12)
        b.predict(u, k * 1.0)
        assert lla == llb
        assert np.array_equal(a.particles, b.particles) and np.array_equal(a.weights, b.weights)
        assert a.index == k + 2


def test_missing_measurement_skips_weight_update():   # Q12, PFtypes.jl:109
    s = lg_model()
    of = s.oracle_filter(200, seed=1)
    of.reset(1)
    of.update(np.zeros(2), np.array([0.3, -0.2]), 0.0)
    w = of.weights.copy()
    ll = of.correct(np.zeros(2), np.array([np.nan, 0.0]), 1.0)
    assert abs(ll) < 1e-12 and np.allclose(of.weights, w, atol=1e-13)


def test_time_index_conventions():   # Q7: forward_trajectory t=(k-1)Ts, loglik t=k*Ts
    q = quadtank_model(t_switch=5.0, a1_factor=2.0)
    T = 12
    u = q.inputs(T)
    of = q.oracle_filter(400, seed=5)
    _, y = of.simulate(u, 1)
    ft = of.forward_trajectory(u, y, epoch=1)
    lk = of.loglik(u, y, epoch=1)
    # identical until the switch can matter (rk4 end time crosses 5 one step earlier for loglik)
    assert np.allclose(ft["ll_steps"][:4], lk["ll_steps"][:4], rtol=0, atol=0)
    assert ft["ll"] != lk["ll"]
    q2 = quadtank_model()   # time-invariant: the two drivers agree exactly
    of2 = q2.oracle_filter(400, seed=5)
    assert of2.forward_trajectory(u, y, epoch=1)["ll"] == of2.loglik(u, y, epoch=1)["ll"]


def test_aux_filter_quirks():   # Q8
    s = lg_model(nx=2, nu=1, ny=1, seed=8)
    N, T = 500, 30
    rng = np.random.default_rng(0)
    u = rng.standard_normal((T, 1))
    of = s.oracle_filter(N, filter=2, seed=3)
    _, y = of.simulate(u, 2)
    ft = of.forward_trajectory(u, y, epoch=1)
    assert abs(ft["ll_steps"][0]) < 1e-12           # first increment is log(sum(1/N)) = 0
    y2 = y.copy(); y2[0] += 5.0                      # y[1] is never weighed in
    assert of.forward_trajectory(u, y2, epoch=1)["ll"] == ft["ll"]
    assert np.all(ft["resampled"][:-1] == 1) and ft["resampled"][-1] == 0   # always resamples, no predict at T
    lk = of.loglik(u, y, epoch=1)
    assert np.allclose(lk["ll_steps"][:-1], ft["ll_steps"][:-1], rtol=0, atol=0)
    assert lk["ll_steps"][-1] != ft["ll_steps"][-1]  # loglik's tail is the inner PF's update!
    # Q9: APF over AdvancedParticleFilter resets the weights -> every increment is ~0
    ofa = s.oracle_filter(N, filter=3, seed=3)
    assert np.all(np.abs(ofa.forward_trajectory(u, y, epoch=1)["ll_steps"]) < 1e-10)


# ---- test/runtests.jl:412-449: PF and APF log-likelihood against the Kalman filter ----------------
def test_pf_and_apf_loglik_vs_kalman():
    A, B, C_ = ref_model_2state()
    n, T, N = 2, 1000, 1000
    rng = np.random.default_rng(0)
    mu0 = rng.standard_normal(n)
    u = rng.standard_normal((T, 1))

    def model(sig):
        return O.ModelArrays(2, 1, 1, C_, sig ** 2 * np.eye(n), np.eye(1), mu0, 4.0 * np.eye(n), A=A, B=B)

    gen = O.OracleFilter(model(0.1), 10, seed=1)     # data from the s = 0.1 model (kf with 0.01*I)
    _, y = gen.simulate(u, 7)
    svec = 10 ** np.linspace(-2, 0, 11)
    llpf = np.array([O.OracleFilter(model(sg), N, filter=0, seed=5).loglik(u, y)["ll"] for sg in svec])
    llapf = np.array([O.OracleFilter(model(sg), N, filter=2, seed=5).loglik(u, y)["ll"] for sg in svec])
    llkf = np.array([O.kalman_loglik(model(sg), u, y) for sg in svec])
    assert 4 <= np.argmax(llkf) <= 6                 # 1-based 5..7
    assert 4 <= np.argmax(llpf) <= 6
    assert 4 <= np.argmax(llapf) <= 6
    assert np.max(np.abs(llkf - llpf)) < 20
    assert np.max(np.abs(llkf - llapf)) < 20


def test_pf_converges_to_kalman_on_config2_model():
    s = lg_model(nx=4, nu=2, ny=2, seed=0)
    T = 60
    u = np.random.default_rng(1).standard_normal((T, 2))
    of = s.oracle_filter(20000, seed=9)
    _, y = of.simulate(u, 4)
    kf = O.kalman_loglik(s.oracle_model(), u, y)
    ll = of.loglik(u, y)["ll"]
    assert abs(ll - kf) < 0.5


@pytest.mark.parametrize("resampling", [1, 2])
def test_stratified_and_residual_in_loop_vs_kalman(resampling):
    """resampling_strategy = ResampleStratified / ResampleResidual inside predict! (resample.jl:12-15 dispatch):
    the filter is still consistent with the closed-form Kalman log-likelihood."""
    s = lg_model(nx=4, nu=2, ny=2, seed=0)
    T = 60
    u = np.random.default_rng(1).standard_normal((T, 2))
    of = s.oracle_filter(8000, seed=9, resampling=resampling)
    _, y = of.simulate(u, 4)
    kf = O.kalman_loglik(s.oracle_model(), u, y)
    r = of.loglik(u, y)
    assert r["resampled"].sum() > 3
    assert abs(r["ll"] - kf) < 0.8
    j = of.ancestors
    assert j.min() >= 1 and j.max() <= 8000


def test_advanced_filter_tracks_state():   # test/runtests.jl:553-599, error bound < 5
    A = np.array([[0.99, 0.1], [0, 0.2]])
    rng = np.random.default_rng(0)
    B = rng.standard_normal((2, 2))
    m = O.ModelArrays(2, 2, 2, np.eye(2), 0.1 ** 2 * np.eye(2), np.eye(2), rng.standard_normal(2), 4 * np.eye(2), A=A, B=B)
    T, N = 200, 500
    u = rng.standard_normal((T, 2))
    of = O.OracleFilter(m, N, filter=1, resample_threshold=0.5, seed=2)
    x, y = of.simulate(u, 3)
    ft = of.forward_trajectory(u, y)
    assert np.linalg.norm(np.mean(x - ft["xhat"], axis=0)) < 5
    assert np.mean((x - ft["xhat"]) ** 2) < 1.0


# ---------------------------------------------------------------------------------------------
# Float32-particle mode of the oracle (config 5): consistency with the f64 restatement
# ---------------------------------------------------------------------------------------------
def test_oracle_f32_mode_tracks_f64_mode():
    """Same model, same RNG streams: the Float32 restatement (ours: summation order of llpf_wide.cuh) must agree
    with the Float64 restatement (pinned to the reference's known answers) to Float32 accuracy over a few steps
    without resampling, and its particles must be exactly representable in Float32."""
    from models import lg_large_model
    s32 = lg_large_model(12, 2, 7, seed=4)
    s64 = lg_large_model(12, 2, 7, seed=4, dtype=np.float64)
    N = 300
    f32 = s32.oracle_filter(N, seed=5, resample_threshold=0.0)
    f64 = s64.oracle_filter(N, seed=5, resample_threshold=0.0)
    f32.reset(1); f64.reset(1)
    assert np.array_equal(f32.particles.astype(np.float32).astype(np.float64), f32.particles)
    assert np.abs(f32.particles - f64.particles).max() < 1e-5
    rng = np.random.default_rng(0)
    for k in range(4):
        u = rng.standard_normal(2)
        y = f64.particles[0] @ s64.C.T + rng.standard_normal(7)
        l32, l64 = f32.correct(u, y, float(k)), f64.correct(u, y, float(k))
        assert abs(l32 - l64) < 2e-4 * max(1.0, abs(l64))
        f32.predict(u, float(k)); f64.predict(u, float(k))
        assert np.abs(f32.particles - f64.particles).max() < 1e-4 * max(1.0, np.abs(f64.particles).max())
        assert np.array_equal(f32.particles.astype(np.float32).astype(np.float64), f32.particles)


def test_oracle_f32_loglik_vs_kalman():
    """PF log-likelihood with Float32 particles against the closed-form Kalman filter (the reference's statistical
    pin, test/runtests.jl:412-449, here for the Float32 restatement): within a few nats at N=4000."""
    from models import lg_large_model
    s = lg_large_model(6, 1, 3, seed=8)
    T = 40
    u = np.random.default_rng(2).standard_normal((T, 1))
    of = s.oracle_filter(4000, seed=3)
    _, y = of.simulate(u, 9)
    ll = of.loglik(u, y, epoch=1)["ll"]
    kf = O.kalman_loglik(s.oracle_model(), u, y)
    assert abs(ll - kf) < 8.0, (ll, kf)


# ---- particle smoother (FFBS)  smoothing.jl:104-143 ; test/runtests.jl:264-333 ----------------------------------
def _py_smooth(of, M, u, xf, wf, wef, uni0, uni):
    """Independent NumPy restatement of smoothing.jl:116-143 + resample.jl:128-152 (systematic initial resample)."""
    T, N, nx = xf.shape
    m = of.model
    L1 = np.linalg.cholesky(np.asarray(m.R1).reshape(nx, nx))
    c0 = -(nx * np.log(2 * np.pi) + 2 * np.log(np.diag(L1)).sum()) / 2
    j, _ = O.resample_systematic(wef[T - 1], uni0, M, j0=np.zeros(M, dtype=np.int64))
    xb = np.zeros((T, M, nx))
    xb[T - 1] = xf[T - 1][j - 1]
    for t in range(T - 1, 0, -1):                      # reference's 1-based t
        fx = np.array([of.dynamics(xf[t - 1, n], u[t - 1], (t - 1) * 1.0) for n in range(N)])
        for mm in range(M):
            r = xb[t, mm][None, :] - fx
            v = np.linalg.solve(L1, r.T).T
            wb = wf[t - 1] + c0 - 0.5 * (v * v).sum(axis=1)
            _, _, we = O.logsumexp(wb)
            bins = np.cumsum(we)
            s = uni[t, mm] * bins[-1]
            i = int(np.searchsorted(bins, s, side="left"))
            xb[t - 1, mm] = xf[t - 1, min(i, N - 1)]
    return xb


def test_smoother_matches_numpy_restatement():
    s = lg_model(nx=2, nu=1, ny=1, seed=3)
    N, T, M = 60, 12, 7
    u = np.random.default_rng(0).standard_normal((T, 1))
    of = s.oracle_filter(N, seed=4)
    _, y = of.simulate(u, 5)
    sol = of.forward_trajectory(u, y, epoch=1, history=True)
    xb = of.smooth(M, u, sol["x"], sol["w"], sol["we"], epoch=1)
    uni0 = O.uniform53(4, 1, 6, 0, 0)
    uni = np.array([[O.uniform53(4, 1, 6, t, mm) for mm in range(M)] for t in range(T)])
    ref = _py_smooth(of, M, u, sol["x"], sol["w"], sol["we"], uni0, uni)
    assert np.array_equal(xb, ref)
    # every smoothed state is one of the filtered particles of its time step
    for t in range(T):
        for mm in range(M):
            assert np.any(np.all(sol["x"][t] == xb[t, mm], axis=1))


def test_smoother_improves_on_filter_mean():   # the (disabled) expectation at test/runtests.jl:346, on an easy model
    s = lg_model(nx=2, nu=1, ny=1, seed=3)
    N, T, M = 300, 60, 40
    u = np.random.default_rng(1).standard_normal((T, 1))
    of = s.oracle_filter(N, seed=2)
    xs, y = of.simulate(u, 8)
    sol = of.forward_trajectory(u, y, epoch=1, history=True)
    xb = of.smooth(M, u, sol["x"], sol["w"], sol["we"], epoch=1)
    xf_mean = np.einsum("tnd,tn->td", sol["x"], sol["we"])
    err_f = np.mean((xf_mean - xs) ** 2)
    err_s = np.mean((xb.mean(axis=1) - xs) ** 2)
    assert err_s < 1.1 * err_f


# ---- user closures (PFtypes.jl:128,232): callbacks must reproduce the descriptor model they restate ------------------
def test_user_function_callbacks_reproduce_descriptor_model():
    s = lg_model(nx=3, nu=2, ny=2, seed=4)
    N, T = 200, 15
    u = np.random.default_rng(2).standard_normal((T, 2))
    ref = s.oracle_filter(N, seed=6)
    _, y = ref.simulate(u, 3)
    a = ref.loglik(u, y, epoch=1)
    L2 = np.linalg.cholesky(s.R2)
    c0 = -(2 * np.log(2 * np.pi) + 2 * np.log(np.diag(L2)).sum()) / 2
    of = s.oracle_filter(N, seed=6)
    of.set_user_functions(dynamics=lambda x, uu, t: s.A @ x + s.B @ uu,
                          loglik=lambda x, uu, yy, t: c0 - 0.5 * np.sum(np.linalg.solve(L2, yy - s.C @ x) ** 2))
    b = of.loglik(u, y, epoch=1)
    assert abs(a["ll"] - b["ll"]) <= 1e-10 * abs(a["ll"])
    assert np.array_equal(a["resampled"], b["resampled"])
    assert np.allclose(ref.particles, of.particles, rtol=0, atol=1e-10)
    # a genuinely nonlinear model runs too (no descriptor could express it)
    of2 = s.oracle_filter(N, seed=6)
    of2.set_user_functions(dynamics=lambda x, uu, t: np.tanh(s.A @ x) + s.B @ uu + 0.01 * t)
    c = of2.loglik(u, y, epoch=1)
    assert np.isfinite(c["ll"]) and c["ll"] != a["ll"]

