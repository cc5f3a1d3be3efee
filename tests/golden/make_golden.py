"""Generates tests/golden/golden_v1.npz from the CPU oracle (oracle/llpf_oracle.c).

The reference (LowLevelParticleFilters.jl) is pure Julia and there is no `julia` in the build image, so these vectors
are NOT outputs of the reference itself: they freeze the oracle restatement (which is pinned against the reference's own
known-answer tests in tests/test_oracle_kat.py) so that neither the oracle nor the CUDA path can drift unnoticed.
Trajectory-level parity with the reference stays unpinned.

    python tests/golden/make_golden.py          # rewrites golden_v1.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from models import lg_large_model, lg_model, quadtank_model  # noqa: E402
from oracle import oracle as O  # noqa: E402


def cases():
    """name -> (spec factory, filter kind, oracle kwargs, N, T, data seed)"""
    return {
        "pf_lg4": (lambda: lg_model(4, 2, 2, seed=0), 0, dict(seed=11), 300, 30, 5),
        "pf_lg2_strat": (lambda: lg_model(2, 1, 1, seed=3), 0, dict(seed=12, resampling=1, resample_threshold=0.5), 257, 25, 6),
        "pf_lg3_resid": (lambda: lg_model(3, 2, 2, seed=1), 0, dict(seed=13, resampling=2, resample_threshold=0.5), 200, 25, 7),
        "apf_lg4": (lambda: lg_model(4, 2, 2, seed=0), 2, dict(seed=14), 300, 20, 8),
        "adv_quadtank": (lambda: quadtank_model(t_switch=10.0, a1_factor=2.0), 1, dict(seed=15), 256, 20, 9),
        "pf_wide_f32": (lambda: lg_large_model(seed=2), 0, dict(seed=16, resample_threshold=0.5), 128, 6, 10),
    }


def run_case(name):
    mk, filt, kw, N, T, dseed = cases()[name]
    s = mk()
    if name == "adv_quadtank":
        u = s.inputs(T)
    else:
        u = np.random.default_rng(dseed).standard_normal((T, s.nu))
    gen = s.oracle_filter(32, seed=1) if name != "adv_quadtank" else s.oracle_filter(32, seed=1)
    _, y = gen.simulate(u, dseed + 100)
    of = s.oracle_filter(N, filter=filt, **kw)
    ft = of.forward_trajectory(u, y, epoch=3, history=(name != "pf_wide_f32"))
    out = {f"{name}/u": u, f"{name}/y": y, f"{name}/ft_ll": ft["ll"], f"{name}/ft_ll_steps": ft["ll_steps"],
           f"{name}/ft_resampled": ft["resampled"], f"{name}/x_final": of.particles.copy(),
           f"{name}/w_final": of.weights.copy(), f"{name}/j_final": of.ancestors.copy()}
    if ft["x"] is not None:
        out[f"{name}/x_hist_first"] = ft["x"][0].copy()
        out[f"{name}/x_hist_last"] = ft["x"][-1].copy()
        out[f"{name}/we_hist_last"] = ft["we"][-1].copy()
    lk = of.loglik(u, y, epoch=4)
    out[f"{name}/loglik"] = lk["ll"]
    out[f"{name}/loglik_resampled"] = lk["resampled"]
    if name == "pf_lg4":
        ft2 = of.forward_trajectory(u, y, epoch=3, history=True)
        out[f"{name}/smooth_xb"] = of.smooth(16, u, ft2["x"], ft2["w"], ft2["we"], epoch=3)
    return out


def resampling_vectors():
    rng = np.random.default_rng(99)
    out = {}
    for N, M in ((10, 10), (257, 257), (100, 37), (64, 200)):
        _, _, we = O.logsumexp(rng.standard_normal(N) * 2)
        u1, uM = rng.random(), rng.random(M)
        j0 = np.full(M, -7, dtype=np.int64)
        js, bs = O.resample_systematic(we, u1, M, j0=j0)
        jt, bt = O.resample_stratified(we, uM, M, j0=j0)
        jr, br = O.resample_residual(we, uM, M, j0=j0, return_bins=True)
        k = f"resample_{N}_{M}"
        out.update({f"{k}/we": we, f"{k}/u1": u1, f"{k}/uM": uM, f"{k}/j_sys": js, f"{k}/bins_sys": bs,
                    f"{k}/j_strat": jt, f"{k}/bins_strat": bt, f"{k}/j_resid": jr, f"{k}/bins_resid": br})
    w = rng.standard_normal(1000) * 5
    ll, wn, we = O.logsumexp(w)
    out.update({"logsumexp/w": w, "logsumexp/ll": ll, "logsumexp/wn": wn, "logsumexp/we": we})
    return out


def build():
    out = {}
    for name in cases():
        out.update(run_case(name))
    out.update(resampling_vectors())
    return out


if __name__ == "__main__":
    d = build()
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **d)
    print(f"wrote {len(d)} arrays, {os.path.getsize(os.path.join(HERE, 'golden_v1.npz')) / 1024:.0f} KiB")
