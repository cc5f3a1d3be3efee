"""Committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from the CPU oracle).
CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path, called through the C-ABI, reproduces them —
bit-exact indices / resample decisions, log-likelihood to 1e-10 relative (bar: 1e-6), particles to 1e-9."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))


def test_oracle_reproduces_golden_vectors():
    d = G.build()
    assert sorted(d) == sorted(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(np.asarray(d[k]), GOLD[k]), k


def _gpu_filter(L, name):
    mk, filt, kw, N, T, _ = G.cases()[name]
    s = mk()
    kw = dict(kw)
    strat = [L.ResampleSystematic, L.ResampleStratified, L.ResampleResidual][kw.pop("resampling", 0)]
    kw["resampling_strategy"] = strat
    if name == "adv_quadtank":
        kw.setdefault("resample_threshold", 0.5)
        pf = s.advanced_filter(N, scan_mode="serial", **kw)
    else:
        pf = s.particle_filter(N, scan_mode="serial", **kw)
    if filt == 2:
        pf = L.AuxiliaryParticleFilter(pf)
    return pf


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.cases()))
def test_gpu_reproduces_golden_trajectories(gpu, name):
    L = gpu
    pf = _gpu_filter(L, name)
    u, y = GOLD[f"{name}/u"], GOLD[f"{name}/y"]
    wide = name == "pf_wide_f32"
    rt, xt = (1e-6, 2e-5) if wide else (1e-10, 1e-9)
    sol = L.forward_trajectory(pf, u, y, epoch=3, history=not wide)
    ll = float(GOLD[f"{name}/ft_ll"])
    assert abs(sol.ll - ll) <= rt * abs(ll)
    assert np.array_equal(sol.extra["resampled"], GOLD[f"{name}/ft_resampled"])
    assert np.allclose(sol.extra["ll_steps"], GOLD[f"{name}/ft_ll_steps"], rtol=0, atol=rt * max(1.0, abs(ll)))
    assert np.allclose(L.particles(pf), GOLD[f"{name}/x_final"], rtol=xt, atol=xt)
    assert np.allclose(L.weights(pf), GOLD[f"{name}/w_final"], rtol=0, atol=1e-4 if wide else 1e-9)
    assert np.array_equal(L.ancestors(pf), GOLD[f"{name}/j_final"])
    if not wide:
        # the whole N x T history (x, w, we of ParticleFilteringSolution) against the oracle, every step
        mk, filt, okw, N, T, _ = G.cases()[name]
        ref = mk().oracle_filter(N, filter=filt, **okw).forward_trajectory(u, y, epoch=3, history=True)
        assert np.allclose(sol.x, ref["x"], rtol=0, atol=xt)
        assert np.allclose(sol.w, ref["w"], rtol=0, atol=1e-9)
        assert np.allclose(sol.we, ref["we"], rtol=1e-8, atol=1e-300)
        assert np.allclose(sol.x[0], GOLD[f"{name}/x_hist_first"], rtol=0, atol=xt)
        assert np.allclose(sol.x[-1], GOLD[f"{name}/x_hist_last"], rtol=0, atol=xt)
        assert np.allclose(sol.we[-1], GOLD[f"{name}/we_hist_last"], rtol=1e-8, atol=1e-300)
    got = L.loglik(pf, u, y, epoch=4, details=True)
    assert abs(got["ll"] - float(GOLD[f"{name}/loglik"])) <= rt * abs(float(GOLD[f"{name}/loglik"]))
    assert np.array_equal(got["resampled"], GOLD[f"{name}/loglik_resampled"])
    if name == "pf_lg4":
        xb, _ = L.smooth(pf, 16, u, y, epoch=3)
        ref = GOLD[f"{name}/smooth_xb"]
        assert np.mean(np.any(np.abs(xb - ref) > 1e-9, axis=2)) <= 0.01


@pytest.mark.gpu
def test_gpu_reproduces_golden_resampling_and_logsumexp(gpu):
    L = gpu
    for N, M in ((10, 10), (257, 257), (100, 37), (64, 200)):
        k = f"resample_{N}_{M}"
        we, u1, uM = GOLD[f"{k}/we"], float(GOLD[f"{k}/u1"]), GOLD[f"{k}/uM"]
        j0 = np.full(M, -7, dtype=np.int64)
        for strat, u, tag in ((L.ResampleSystematic, u1, "sys"), (L.ResampleStratified, uM, "strat"),
                              (L.ResampleResidual, uM, "resid")):
            j, b = L.resample(strat, we, u, M, j0=j0, scan_mode="serial", return_bins=True)
            assert np.array_equal(j, GOLD[f"{k}/j_{tag}"]), (k, tag)
            assert np.array_equal(b, GOLD[f"{k}/bins_{tag}"]), (k, tag)
    ll, wn, we = L.logsumexp(GOLD["logsumexp/w"])
    assert abs(ll - float(GOLD["logsumexp/ll"])) <= 1e-13 * abs(float(GOLD["logsumexp/ll"]))
    assert np.allclose(wn, GOLD["logsumexp/wn"], rtol=0, atol=1e-12)
    assert np.allclose(we, GOLD["logsumexp/we"], rtol=1e-12, atol=1e-300)
