"""Second-source check of the oracle: oracle/pyref.py — a pure-Python restatement of the
reference's PF / AdvancedPF / APF `forward_trajectory` and `loglik`, written from the Julia sources — must reproduce the
C oracle (oracle/llpf_oracle.c) BIT FOR BIT on the golden cases: particles, log-weights, exp-weights of every step,
ancestors, resample decisions, per-step and total log-likelihood."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import pyref as P  # noqa: E402


def _py_model(s, name):
    if name == "adv_quadtank":
        return P.Model(s.C, s.R1, s.R2, s.x0, s.R1,
                       quadtank=dict(p=s.p, t_switch=s.t_switch, a1_factor=s.a1_factor, Ts=s.Ts, supersample=s.supersample))
    return P.Model(s.C, s.R1, s.R2, s.mu0, s.Sigma0, A=s.A, B=s.B)


CASES = [n for n in G.cases() if n != "pf_wide_f32"]      # the Float32 mode's summation order is ours, not the reference's


@pytest.mark.parametrize("name", CASES)
def test_python_twin_reproduces_c_oracle_bit_for_bit(name):
    mk, filt, kw, N, T, dseed = G.cases()[name]
    s = mk()
    N, T = min(N, 200), min(T, 20)          # pure-Python loops
    u = s.inputs(T) if name == "adv_quadtank" else np.random.default_rng(dseed).standard_normal((T, s.nu))
    _, y = s.oracle_filter(32, seed=1).simulate(u, dseed + 100)
    of = s.oracle_filter(N, filter=filt, **kw)
    ref = of.forward_trajectory(u, y, epoch=3, history=True)
    okw = dict(kw)
    thr = okw.get("resample_threshold", 0.5 if name == "adv_quadtank" else 0.1)
    pf = P.Filter(_py_model(s, name), N, kind=filt, resampling=okw.get("resampling", 0), resample_threshold=thr,
                  seed=okw["seed"])
    ul, yl = [list(r) for r in u], [list(r) for r in y]
    got = pf.forward_trajectory(ul, yl, epoch=3)
    assert got["ll"] == ref["ll"]
    assert got["ll_steps"] == list(ref["ll_steps"])
    assert got["resampled"] == list(ref["resampled"])
    assert sum(got["resampled"]) > 0
    assert np.array_equal(np.array(got["x"]), ref["x"])
    assert np.array_equal(np.array(got["w"]), ref["w"])
    assert np.array_equal(np.array(got["we"]), ref["we"])
    assert pf.j == list(of.ancestors)
    assert np.array_equal(np.array(pf.x), of.particles)
    # loglik: the other time convention (t = index*Ts), APF inner-filter tail
    lk = of.loglik(u, y, epoch=4)
    g2 = pf.loglik(ul, yl, epoch=4)
    assert g2["ll"] == lk["ll"]
    assert g2["resampled"] == list(lk["resampled"])
    assert np.array_equal(np.array(pf.x), of.particles)
    assert np.array_equal(np.array(pf.w), of.weights)


def test_python_twin_aux_over_advanced_and_missing_measurements():
    """APF{AdvancedPF} (filtering.jl:219-234: ll == 0 after the first step) and `missing` measurements (PFtypes.jl:109)."""
    from models import lg_model
    s = lg_model(2, 1, 1, seed=5)
    N, T = 64, 12
    u = np.random.default_rng(2).standard_normal((T, 1))
    _, y = s.oracle_filter(16, seed=1).simulate(u, 9)
    y[3] = np.nan
    for filt in (3, 0):
        of = s.oracle_filter(N, filter=filt, seed=21, resample_threshold=0.5)
        ref = of.forward_trajectory(u, y, epoch=1, history=True)
        pf = P.Filter(P.Model(s.C, s.R1, s.R2, s.mu0, s.Sigma0, A=s.A, B=s.B), N, kind=filt, resample_threshold=0.5, seed=21)
        got = pf.forward_trajectory([list(r) for r in u], [list(r) for r in y], epoch=1)
        assert got["ll_steps"] == list(ref["ll_steps"])
        assert np.array_equal(np.array(got["x"]), ref["x"])
        assert np.array_equal(np.array(got["we"]), ref["we"])
        if filt == 3:
            assert all(v == 0.0 for v in got["ll_steps"][1:])


def test_python_twin_rng_and_resamplers_match_c_oracle():
    assert list(P.philox4x32_10((0, 0, 0, 0), (0, 0))) == O.philox4x32_10((0, 0, 0, 0), (0, 0))
    assert list(P.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2)) == O.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2)
    for (seed, ep, st, t, i) in [(0, 0, 0, 0, 0), (7, 3, 1, 5, 123456), (2 ** 40 + 5, 2, 1, 77, 2 ** 33 + 9)]:
        assert P.normals(seed, ep, st, t, i, 7) == list(O.normals(seed, ep, st, t, i, 7))
        assert P.uniform53(seed, ep, st, t, i) == O.uniform53(seed, ep, st, t, i)
    rng = np.random.default_rng(0)
    for N, M in ((10, 10), (257, 257), (100, 37), (64, 200)):
        _, _, we = O.logsumexp(rng.standard_normal(N) * 2)
        u1, uM = rng.random(), rng.random(M)
        j0 = [-7] * M
        js, bs = O.resample_systematic(we, u1, M, j0=np.array(j0))
        b = [0.0] * N
        assert P.resample_systematic(list(we), list(j0), b, u1, M) == list(js) and b == list(bs)
        jt, bt = O.resample_stratified(we, uM, M, j0=np.array(j0))
        assert P.resample_stratified(list(we), list(j0), b, list(uM), M) == list(jt) and b == list(bt)
        jr, br = O.resample_residual(we, uM, M, j0=np.array(j0), return_bins=True)
        assert P.resample_residual(list(we), list(j0), b, list(uM), M) == list(jr) and b == list(br)
