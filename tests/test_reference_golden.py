"""Consumer of tests/golden/reference_v1.json — vectors dumped from the REAL LowLevelParticleFilters.jl by
julia/dump_golden.jl (every random variate the reference consumed + everything it computed from them).

The build image has no julia, so the file may be absent: the tests that need it then SKIP with a reason, and
`test_dump_schema_round_trip` exercises the identical consumer code on a dump of the same schema written by the Python
restatement itself (plumbing check, not a parity claim).  The moment someone runs

    julia julia/dump_golden.jl          # writes tests/golden/reference_v1.json

these tests pin the oracle (and through tests/test_pyref_twin.py + the GPU parity tests, the CUDA path) to the reference."""
import json
import math
import os

import numpy as np
import pytest

from oracle import julia_range as J
from oracle import pyref as P

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "reference_v1.json")
STRAT = {"systematic": 0, "stratified": 1, "residual": 2}


def _model(c):
    return P.Model(c["C"], c["R1"], c["R2"], c["mu0"], c["Sigma0"], A=c["A"], B=c["B"])


def _close(a, b, rtol=1e-12, atol=1e-13):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def check_case(c, exact=False):
    """Feed the recorded variates to the restatement and compare with what the reference computed."""
    kind = P.AUX if c["filter"] == "apf" else P.PF
    for mode in ("forward_trajectory", "loglik"):
        run = c[mode]
        pf = P.Filter(_model(c), c["N"], kind=kind, resampling=STRAT[c["resampling"]], resample_threshold=c["threshold"],
                      Ts=c["Ts"], inject=dict(x0=run["x0"], noise=run["noise"], u_res=run["u_res"]))
        # drive step by step exactly like julia/dump_golden.jl `recorded_run`, recording the state after every correct!
        pf.reset(0)
        T = c["T"]
        u, y = c["u"], c["y"]
        xs, ws, wes, lls, res = [], [], [], [], []
        for t in range(1, T + 1):
            aux_tail = kind == P.AUX and mode == "loglik" and t == T
            if mode == "forward_trajectory" or kind == P.AUX:
                ti = (t - 1) * pf.Ts
            else:
                ti = pf.t * pf.Ts
            if aux_tail:
                pf.kind = P.PF
            lls.append(pf.correct(u[t - 1], y[t - 1], ti))
            xs.append([list(v) for v in pf.x]); ws.append(list(pf.w)); wes.append(list(pf.we))
            n0 = pf.nres
            if kind == P.AUX and not aux_tail:
                if t < T:
                    pf.predict_aux(u[t - 1], y[t], ti)
            else:
                pf.predict(u[t - 1], ti)
            res.append(pf.nres - n0)
            pf.kind = kind
        assert res == list(run["resampled"]), (c["name"], mode)
        assert pf.j == list(run["j_final"]), (c["name"], mode)
        if exact:
            assert xs == run["x"] and ws == run["w"] and wes == run["we"] and lls == run["ll_steps"]
        else:
            assert _close(xs, run["x"]), (c["name"], mode)
            assert _close(ws, run["w"], rtol=1e-11, atol=1e-11), (c["name"], mode)
            assert _close(wes, run["we"], rtol=1e-10, atol=1e-300), (c["name"], mode)
            assert _close(lls, run["ll_steps"], rtol=1e-11, atol=1e-11), (c["name"], mode)
            assert abs(sum(lls) - run["ll"]) <= 1e-10 * max(1.0, abs(run["ll"]))
        assert _close([list(v) for v in pf.x], run["x_final"])


def check_ranges(doc):
    for r in doc.get("ranges", []):
        s = J.julia_thresholds(r["r"], r["M"], r["total"])
        assert (s.ref_hi, s.ref_lo, s.step_hi, s.step_lo, s.offset, s.len) == \
               (r["ref_hi"], r["ref_lo"], r["step_hi"], r["step_lo"], r["offset"], r["len"]), r
        assert [s.getindex(i) for i in range(1, r["M"] + 1)] == r["s"]


def check_resample(doc):
    for r in doc.get("resample", []):
        N, M = r["N"], r["M"]
        j, b = [-7] * M, [0.0] * N
        if "Systematic" in r["strategy"]:
            P.resample_systematic(list(r["we"]), j, b, r["draws"][0], M)
        elif "Stratified" in r["strategy"]:
            P.resample_stratified(list(r["we"]), j, b, r["draws"], M)
        else:
            P.resample_residual(list(r["we"]), j, b, r["draws"], M)
        assert j == list(r["j"]), r["strategy"]


def check_rbpf(c, exact=False):
    """`rbpf` section: the mixed model of test/test_rbpf.jl (fn = xn, g = xn, An != 0) driven like forward_trajectory with the
    variates the reference consumed; state after every correct! as flat particles [xn; xl; tril(R)]."""
    from oracle import rbpf_ref as R
    kf = dict(A=c["A"], B=[[] for _ in c["A"]], C=c["C"], R1=c["R1l"], R2=c["R2"], mu0=c["x0l"], Sigma0=c["R0"])
    pf = R.RBPFRef(c["N"], kf, lambda xn, u, t: list(xn), lambda xn, u, t: list(xn), c["R1n"], (c["x0n"], c["R1n"]),
                   An=c["An"], resample_threshold=c["threshold"], Ts=c["Ts"],
                   inject=dict(x0=c["x0"], noise=c["noise"], u_res=c["u_res"]))
    pf.reset(0)
    xs, ws, wes, lls, res = [], [], [], [], []
    for t in range(1, c["T"] + 1):
        ti = (t - 1) * pf.Ts
        lls.append(pf.correct([], c["y"][t - 1], ti))
        xs.append(pf.particles_flat()); ws.append(list(pf.w)); wes.append(list(pf.we))
        n0 = pf.nres
        pf.predict([], ti)
        res.append(pf.nres - n0)
    assert res == list(c["resampled"]) and pf.j == list(c["j_final"])
    if exact:
        assert xs == c["x"] and ws == c["w"] and wes == c["we"] and lls == c["ll_steps"]
    else:
        assert _close(xs, c["x"], rtol=1e-11, atol=1e-12) and _close(ws, c["w"], rtol=1e-10, atol=1e-10)
        assert _close(wes, c["we"], rtol=1e-9, atol=1e-300) and _close(lls, c["ll_steps"], rtol=1e-10, atol=1e-10)
    assert _close(pf.particles_flat(), c["x_final"], rtol=1e-11, atol=1e-12)


def _python_rbpf_dump():
    """the `rbpf` section with the schema of julia/dump_golden.jl, written by the restatement with its own variates"""
    from oracle import rbpf_ref as R
    rng = np.random.default_rng(5)
    T, N = 25, 80
    xn, xl, y = 1.0, 1.0, []
    for _ in range(T):
        y.append([xn + xl + math.sqrt(0.1) * rng.standard_normal()])
        xn, xl = xn + 0.5 * xl + 0.1 * rng.standard_normal(), 0.95 * xl + 0.1 * rng.standard_normal()
    c = dict(name="rbpf_mixed", N=N, T=T, threshold=0.5, Ts=1.0, A=[[0.95]], An=[[0.5]], C=[[1.0]], R1l=[[0.01]], R1n=[[0.01]],
             R2=[[0.1]], x0n=[1.0], x0l=[1.0], R0=[[1.0]], y=y)
    rec = {}
    kf = dict(A=c["A"], B=[[]], C=c["C"], R1=c["R1l"], R2=c["R2"], mu0=c["x0l"], Sigma0=c["R0"])
    pf = R.RBPFRef(N, kf, lambda xn, u, t: list(xn), lambda xn, u, t: list(xn), c["R1n"], (c["x0n"], c["R1n"]), An=c["An"],
                   resample_threshold=0.5, seed=3, record=rec)
    pf.reset(1)
    xs, ws, wes, lls, res = [], [], [], [], []
    for t in range(1, T + 1):
        lls.append(pf.correct([], y[t - 1], (t - 1) * 1.0))
        xs.append(pf.particles_flat()); ws.append(list(pf.w)); wes.append(list(pf.we))
        n0 = pf.nres
        pf.predict([], (t - 1) * 1.0)
        res.append(pf.nres - n0)
    c.update(x0=rec["x0"], noise=rec["noise"], u_res=rec["u_res"], x=xs, w=ws, we=wes, ll_steps=lls, ll=sum(lls), resampled=res,
             j_final=list(pf.j), x_final=pf.particles_flat())
    return c


def test_rbpf_dump_schema_round_trip(tmp_path):
    c = _python_rbpf_dump()
    assert sum(c["resampled"]) >= 2
    path = tmp_path / "rbpf.json"
    path.write_text(json.dumps(dict(rbpf=[c])))
    check_rbpf(json.loads(path.read_text())["rbpf"][0], exact=True)


def _enkf_ref(c, **kw):
    from oracle import enkf_ref as E
    A, B = c["A"], c["B"]

    def dyn(x, u, t):                     # dynamics(x,u,p,t) = A*x .+ B*u  (StaticArrays: left-to-right sums)
        return [P.matvec(A, x)[r] + P.matvec(B, u)[r] for r in range(len(x))]
    return E.EnKFRef(dyn, c["C"], c["R1"], c["R2"], c["mu0"], c["Sigma0"], c["N"], Ts=c["Ts"], inflation=c["inflation"], **kw)


def check_enkf(c, exact=False):
    """`enkf` section: forward_trajectory of the ensemble Kalman filter on the recorded standard normals"""
    ref = _enkf_ref(c, inject=dict(z0=c["z0"], zdyn=c["zdyn"], zobs=c["zobs"]))
    out = ref.forward_trajectory(c["u"], c["y"])
    for key in ("x", "R", "xt", "Rt", "e", "S", "K", "ll_steps"):
        if exact:
            assert out[key] == c[key], key
        else:
            assert _close(out[key], c[key], rtol=1e-10, atol=1e-11), key     # the reference's BLAS / mean() sum in other orders
    assert abs(out["ll"] - c["ll"]) <= 1e-10 * max(1.0, abs(c["ll"]))
    assert _close(ref.X, c["ensemble_final"], rtol=1e-10, atol=1e-11)


def test_enkf_dump_schema_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    T = 12
    c = dict(name="enkf_lin2", N=60, T=T, Ts=1.0, inflation=1.05, A=[[0.99, 0.1], [0.0, 0.2]], B=[[-0.74, 1.61], [-1.44, 1.75]],
             C=[[1.0, 0.0], [0.0, 1.0]], R1=[[1.0, 0.0], [0.0, 1.0]], R2=[[1.0, 0.0], [0.0, 1.0]], mu0=[0.3, -0.2],
             Sigma0=[[4.0, 0.0], [0.0, 4.0]], u=rng.standard_normal((T, 2)).tolist(), y=rng.standard_normal((T, 2)).tolist())
    rec = {}
    ref = _enkf_ref(c, seed=4, record=rec)
    out = ref.forward_trajectory(c["u"], c["y"], epoch=1)
    c.update({k: out[k] for k in ("x", "R", "xt", "Rt", "e", "S", "K", "ll_steps", "ll")})
    c.update(z0=rec["z0"], zdyn=rec["zdyn"], zobs=rec["zobs"], ensemble_final=[list(v) for v in ref.X])
    path = tmp_path / "enkf.json"
    path.write_text(json.dumps(dict(enkf=[c])))
    check_enkf(json.loads(path.read_text())["enkf"][0], exact=True)


def _python_dump():
    """A dump with the schema of julia/dump_golden.jl, produced by the Python restatement with its own (Philox) variates."""
    from models import lg_model
    cases = []
    for name, filt, strat, thr, nx, nu, ny, N, T in (("pf_sys", "pf", "systematic", 0.5, 4, 2, 2, 60, 12),
                                                    ("pf_resid", "pf", "residual", 0.5, 3, 2, 2, 48, 10),
                                                    ("apf_sys", "apf", "systematic", 0.1, 2, 1, 1, 50, 9)):
        s = lg_model(nx, nu, ny, seed=3)
        u = np.random.default_rng(1).standard_normal((T, nu))
        _, y = s.oracle_filter(16, seed=1).simulate(u, 5)
        c = dict(name=name, filter=filt, resampling=strat, threshold=thr, N=N, T=T, nx=nx, nu=nu, ny=ny, Ts=1.0,
                 A=s.A.tolist(), B=s.B.tolist(), C=s.C.tolist(), R1=s.R1.tolist(), R2=s.R2.tolist(), mu0=s.mu0.tolist(),
                 Sigma0=s.Sigma0.tolist(), u=u.tolist(), y=y.tolist())
        kind = P.AUX if filt == "apf" else P.PF
        for mode in ("forward_trajectory", "loglik"):
            rec = {}
            pf = P.Filter(_model(c), N, kind=kind, resampling=STRAT[strat], resample_threshold=thr, seed=9, record=rec)
            if mode == "forward_trajectory":
                out = pf.forward_trajectory(c["u"], c["y"], epoch=1)
                run = dict(x=out["x"], w=out["w"], we=out["we"], ll_steps=out["ll_steps"], ll=out["ll"],
                           resampled=out["resampled"])
            else:
                # loglik records no per-step state: rebuild it with a second, injected pass
                out = pf.loglik(c["u"], c["y"], epoch=1)
                run = dict(ll=out["ll"], resampled=out["resampled"])
            run.update(x0=rec["x0"], noise=rec["noise"], u_res=rec["u_res"], j_final=list(pf.j),
                       x_final=[list(v) for v in pf.x], w_final=list(pf.w))
            c[mode] = run
        cases.append(c)
    return dict(version=1, julia="python-emulation", package="pyref", cases=cases)


def test_dump_schema_round_trip(tmp_path):
    doc = _python_dump()
    path = tmp_path / "reference_v1.json"
    path.write_text(json.dumps(doc))
    doc2 = json.loads(path.read_text())
    for c in doc2["cases"]:
        ft = c["forward_trajectory"]
        # forward_trajectory: every recorded quantity must come back bit for bit through the injected variates
        kind = P.AUX if c["filter"] == "apf" else P.PF
        pf = P.Filter(_model(c), c["N"], kind=kind, resampling=STRAT[c["resampling"]], resample_threshold=c["threshold"],
                      inject=dict(x0=ft["x0"], noise=ft["noise"], u_res=ft["u_res"]))
        out = pf.forward_trajectory(c["u"], c["y"])
        assert out["x"] == ft["x"] and out["w"] == ft["w"] and out["we"] == ft["we"] and out["ll"] == ft["ll"]
        lk = c["loglik"]
        pf = P.Filter(_model(c), c["N"], kind=kind, resampling=STRAT[c["resampling"]], resample_threshold=c["threshold"],
                      inject=dict(x0=lk["x0"], noise=lk["noise"], u_res=lk["u_res"]))
        out = pf.loglik(c["u"], c["y"])
        assert out["ll"] == lk["ll"] and out["resampled"] == lk["resampled"] and pf.j == lk["j_final"]
        # and the stepwise consumer used for the real file agrees with the drivers
        full = dict(c)
        pf2 = P.Filter(_model(c), c["N"], kind=kind, resampling=STRAT[c["resampling"]], resample_threshold=c["threshold"],
                       inject=dict(x0=lk["x0"], noise=lk["noise"], u_res=lk["u_res"]))
        full["loglik"] = dict(lk, **_stepwise_state(pf2, c, "loglik"))
        check_case(full, exact=True)


def _stepwise_state(pf, c, mode):
    kind = pf.kind
    pf.reset(0)
    xs, ws, wes, lls = [], [], [], []
    T = c["T"]
    for t in range(1, T + 1):
        aux_tail = kind == P.AUX and mode == "loglik" and t == T
        ti = (t - 1) * pf.Ts if (mode == "forward_trajectory" or kind == P.AUX) else pf.t * pf.Ts
        if aux_tail:
            pf.kind = P.PF
        lls.append(pf.correct(c["u"][t - 1], c["y"][t - 1], ti))
        xs.append([list(v) for v in pf.x]); ws.append(list(pf.w)); wes.append(list(pf.we))
        if kind == P.AUX and not aux_tail:
            if t < T:
                pf.predict_aux(c["u"][t - 1], c["y"][t], ti)
        else:
            pf.predict(c["u"][t - 1], ti)
        pf.kind = kind
    return dict(x=xs, w=ws, we=wes, ll_steps=lls)


needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="tests/golden/reference_v1.json absent: run "
                               "`julia julia/dump_golden.jl` on a machine with Julia + LowLevelParticleFilters.jl "
                               "(the build image has neither)")


@needs_ref
def test_restatement_reproduces_the_reference_trajectories():
    doc = json.load(open(REF))
    assert doc["version"] == 1 and len(doc["cases"]) >= 4
    for c in doc["cases"]:
        check_case(c)


@needs_ref
def test_restatement_reproduces_the_reference_rbpf():
    doc = json.load(open(REF))
    if not doc.get("rbpf"):
        pytest.skip("this dump has no `rbpf` section (written by an older julia/dump_golden.jl)")
    for c in doc["rbpf"]:
        check_rbpf(c)


@needs_ref
def test_restatement_reproduces_the_reference_enkf():
    doc = json.load(open(REF))
    if not doc.get("enkf"):
        pytest.skip("this dump has no `enkf` section (written by an older julia/dump_golden.jl)")
    for c in doc["enkf"]:
        check_enkf(c)


@needs_ref
def test_restatement_reproduces_the_reference_ranges_and_resampling():
    doc = json.load(open(REF))
    check_ranges(doc)
    check_resample(doc)
