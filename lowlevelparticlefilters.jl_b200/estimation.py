"""Particle marginal Metropolis-Hastings on top of the device `loglik`.

Mirrors the reference's host-side driver:
    log_likelihood_fun(filter_from_parameters, priors, u, y, p)    src/smoothing.jl:266-283
    naive_sampler(θ₀)                                              src/smoothing.jl:285-288
    metropolis(ll, R, θ₀, draw)                                    src/smoothing.jl:311-328
    metropolis_threaded(burnin, ll, R, θ₀, draw; nthreads)         src/smoothing.jl:335-347
Every `ll(θ)` is one `loglik(pf, u, y)` = one persistent-kernel launch on the GPU (llpf_run); chains of
`metropolis_threaded` run on host threads (ctypes releases the GIL inside the C-ABI call) with one filter handle =
one CUDA stream per chain, so the chains' kernels overlap on the device: a PMMH filter of N ~ 10^3 particles occupies
a handful of thread blocks and ~70 such chains fit a B200 side by side.
"""
import inspect
import math
import threading

import numpy as np

from ._abi import LLPFError
from . import filters as F


class Normal:
    """Normal(mu, sigma) prior with logpdf (Distributions.Normal; used at example_lineargaussian.jl:201)."""

    def __init__(self, mu=0.0, sigma=1.0):
        self.mu, self.sigma = float(mu), float(sigma)

    def logpdf(self, x):
        z = (float(x) - self.mu) / self.sigma
        return -0.5 * z * z - math.log(self.sigma) - 0.5 * math.log(2 * math.pi)


class Uniform:
    def __init__(self, a=0.0, b=1.0):
        self.a, self.b = float(a), float(b)

    def logpdf(self, x):
        return -math.log(self.b - self.a) if self.a <= x <= self.b else -math.inf


def set_model(pf, *, dynamics=None, measurement=None, dynamics_density=None, measurement_density=None,
              initial_density=None, measurement_likelihood=None):
    """Replace model matrices / noise covariances of an existing filter in place (llpf_set_model): what
    `filter_from_parameters(θ, pf)` does when it reuses `pf` (smoothing.jl:277) — no device memory is reallocated.
    Dimensions and the dynamics kind cannot change."""
    import ctypes as C
    tgt = pf
    dyn = dynamics if dynamics is not None else tgt.dynamics
    d1 = dynamics_density if dynamics_density is not None else tgt.dynamics_density
    d0 = initial_density if initial_density is not None else tgt.initial_density
    if isinstance(getattr(tgt, "pf", tgt), F.AdvancedParticleFilter) or isinstance(tgt, F.AdvancedParticleFilter):
        base = getattr(tgt, "pf", tgt)
        lik = measurement_likelihood if measurement_likelihood is not None else base.measurement_likelihood
        Cm, R2 = lik.C, lik.R2
        base.measurement_likelihood = lik
    else:
        base = getattr(tgt, "pf", tgt)
        meas = measurement if measurement is not None else base.measurement
        d2 = measurement_density if measurement_density is not None else base.measurement_density
        Cm, R2 = meas.C, d2.Sigma
        base.measurement, base.measurement_density = meas, d2
    model = F._ModelBuffers(dyn, Cm, d1.Sigma, R2, d0)
    F.check(pf._lib, pf._lib.llpf_set_model(pf._h, C.byref(model.struct)))
    for obj in {id(tgt): tgt, id(base): base}.values():
        obj.dynamics, obj.dynamics_density, obj.initial_density = dyn, d1, d0
    pf._model = model
    return pf


def log_likelihood_fun(filter_from_parameters, priors, u, y, p=None):
    """θ -> log p(y|θ) + log p(θ)   smoothing.jl:266-283.
    `filter_from_parameters(θ)` builds a filter; if it also accepts `(θ, pf)` it is handed the previous filter to
    update in place (use `set_model`), exactly like the reference's two-argument call at :277."""
    try:
        nargs = len(inspect.signature(filter_from_parameters).parameters)
    except (TypeError, ValueError):
        nargs = 1
    # the RNG epoch belongs to the likelihood function, not to the filter handle: a filter_from_parameters that builds a
    # fresh filter per call would otherwise evaluate every θ with the same particle noise (a fixed, biased surface
    # instead of the reference's fresh unbiased estimate per call — pf.rng runs on across reset!, filtering.jl:6)
    state = {"pf": None, "epoch": 0}

    def ll(theta):
        theta = np.asarray(theta, dtype=np.float64).reshape(-1)
        if theta.size != len(priors):
            raise ValueError("Input must have same length as priors")
        lp = sum(priors[i].logpdf(theta[i]) for i in range(len(priors)))
        if not math.isfinite(lp):
            return -math.inf
        try:
            if state["pf"] is None or nargs < 2:
                state["pf"] = filter_from_parameters(theta)
            else:
                state["pf"] = filter_from_parameters(theta, state["pf"])
            state["epoch"] += 1
            return lp + F.loglik(state["pf"], u, y, epoch=state["epoch"] & 0xFFFFFF)
        except (LLPFError, np.linalg.LinAlgError, FloatingPointError):
            return -math.inf     # the reference's `catch` at :280 (e.g. a covariance that is not positive definite)

    ll.state = state
    return ll


def naive_sampler(theta0, rng=None):
    """θ -> θ + N(0, diag(0.1|θ₀|))   smoothing.jl:285-288 (covariance 0.1|θ₀|, i.e. std sqrt(0.1|θ₀|))"""
    theta0 = np.asarray(theta0, dtype=np.float64)
    if np.any(theta0 == 0):
        raise ValueError("Naive sampler does not work if initial parameter vector contains zeros")
    rng = np.random.default_rng() if rng is None else rng
    sd = np.sqrt(0.1 * np.abs(theta0))
    return lambda th: np.asarray(th) + sd * rng.standard_normal(theta0.size)


def metropolis(ll, R, theta0, draw=None, rng=None):
    """params, lls = metropolis(ll, R, θ₀, draw)   smoothing.jl:311-328 (symmetric proposal, R iterations)"""
    rng = np.random.default_rng() if rng is None else rng
    theta0 = np.asarray(theta0, dtype=np.float64)
    draw = naive_sampler(theta0, rng) if draw is None else draw
    params = np.zeros((R, theta0.size))
    lls = np.zeros(R)
    params[0] = theta0
    lls[0] = ll(theta0)
    for i in range(1, R):
        th = np.asarray(draw(params[i - 1]), dtype=np.float64)
        lli = ll(th)
        d = lli - lls[i - 1]
        # rand() < exp(lli - lls[i-1]); nan (both -Inf) rejects, like the reference's comparison with NaN
        if not math.isnan(d) and rng.random() < math.exp(min(d, 0.0)):
            params[i], lls[i] = th, lli
        else:
            params[i], lls[i] = params[i - 1], lls[i - 1]
    return params, lls


def metropolis_threaded(burnin, ll, R, theta0, draw=None, *, nthreads=4, seed=None):
    """metropolis_threaded(burnin, ll, R, θ₀, draw; nthreads)   smoothing.jl:335-347: `nthreads` independent chains,
    returns [(R - burnin) * nthreads] x [len(θ) + 1] with the log-likelihoods in the last column.
    `ll` may be a function (shared by the chains, as in the reference — calls are then serialised by a lock, because one
    filter handle is one mutable state) or a zero-argument FACTORY returning a fresh ll per chain (own filter handle,
    own CUDA stream): the chains' kernels then run concurrently on the GPU."""
    factory = None
    try:
        if len(inspect.signature(ll).parameters) == 0:
            factory = ll
    except (TypeError, ValueError):
        pass
    lock = threading.Lock()
    res = [None] * nthreads
    errs = []
    ss = np.random.SeedSequence(seed)
    kids = ss.spawn(nthreads)

    def run(k):
        try:
            rng = np.random.default_rng(kids[k])
            if factory is not None:
                llk = factory()
                if hasattr(llk, "state"):      # chains must not share particle-noise streams (per-task RNGs in the reference)
                    llk.state["epoch"] = (k + 1) << 18
            else:
                def llk(th):
                    with lock:
                        return ll(th)
            drawk = draw(rng) if (draw is not None and getattr(draw, "wants_rng", False)) else draw
            p, l = metropolis(llk, R, theta0, drawk, rng)
            res[k] = np.hstack([p, l[:, None]])[burnin:]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(k,)) for k in range(nthreads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return np.vstack(res)


def metropolis_batched(burnin, filter_from_parameters, priors, u, y, R, theta0, draw=None, *, nchains=64, seed=None):
    """The device-side form of `metropolis_threaded` (smoothing.jl:335-347): `nchains` independent Markov chains advance in
    lock step and every iteration evaluates all their log-likelihoods with ONE kernel launch (loglik_batch: one thread
    block per chain) instead of one launch per chain per iteration.
    filter_from_parameters(θ) must build a filter with single_block=True; if it accepts (θ, pf) the chain's filter is
    updated in place with set_model (no device allocation per proposal).  Each chain's filter gets its own seed.
    Returns [(R - burnin) * nchains] x [len(θ) + 1] with the log-likelihoods in the last column, like the reference."""
    ss = np.random.SeedSequence(seed)
    rngs = [np.random.default_rng(k) for k in ss.spawn(nchains)]
    theta0 = np.asarray(theta0, dtype=np.float64).reshape(-1)
    try:
        nargs = len(inspect.signature(filter_from_parameters).parameters)
    except (TypeError, ValueError):
        nargs = 1
    draws = [(naive_sampler(theta0, rngs[c]) if draw is None else (draw(rngs[c]) if getattr(draw, "wants_rng", False) else draw))
             for c in range(nchains)]

    def prior(th):
        return sum(priors[i].logpdf(th[i]) for i in range(len(priors)))

    pfs = [None] * nchains

    def evaluate(thetas, it):
        """log p(y|θ_c) + log p(θ_c) for every chain; chains whose proposal has zero prior mass (or an invalid model) get
        -Inf and are evaluated with their previous model (result discarded)."""
        lp = np.array([prior(th) for th in thetas])
        ok = np.isfinite(lp)
        for c in range(nchains):
            if not ok[c] and pfs[c] is not None:
                continue
            try:
                if pfs[c] is None or nargs < 2:
                    pf = filter_from_parameters(thetas[c] if ok[c] else theta0)
                    pf.seed_chain = c
                    pfs[c] = pf
                else:
                    pfs[c] = filter_from_parameters(thetas[c], pfs[c])
            except (LLPFError, np.linalg.LinAlgError, FloatingPointError):
                ok[c] = False
        ll = F.loglik_batch(pfs, u, y, epochs=[((c + 1) << 18) + it + 1 for c in range(nchains)])
        out = np.where(ok & np.isfinite(ll), lp + ll, -math.inf)
        return out

    params = np.zeros((R, nchains, theta0.size))
    lls = np.zeros((R, nchains))
    params[0] = theta0
    lls[0] = evaluate([theta0] * nchains, 0)
    for i in range(1, R):
        prop = [np.asarray(draws[c](params[i - 1, c]), dtype=np.float64) for c in range(nchains)]
        lli = evaluate(prop, i)
        for c in range(nchains):
            d = lli[c] - lls[i - 1, c]
            if not math.isnan(d) and rngs[c].random() < math.exp(min(d, 0.0)):
                params[i, c], lls[i, c] = prop[c], lli[c]
            else:
                params[i, c], lls[i, c] = params[i - 1, c], lls[i - 1, c]
    out = np.concatenate([params, lls[:, :, None]], axis=2)[burnin:]        # [R - burnin][chains][nθ + 1]
    return np.concatenate([out[:, c] for c in range(nchains)], axis=0)


# ---------------------------------------------------------------------------------------------
# weighted statistics of a stored solution   filtering.jl:570-595
# ---------------------------------------------------------------------------------------------
def weighted_cov(x, we=None):
    """weighted_cov(x, we) / weighted_cov(sol)  filtering.jl:575-583: per time step the covariance of the particles with
    ProbabilityWeights(we), corrected=true (StatsBase: factor n/((n-1) sum(w)), n = number of non-zero weights)."""
    if we is None:
        x, we = x.x, x.we
    x, we = np.asarray(x), np.asarray(we)
    out = []
    for t in range(x.shape[0]):
        w = we[t]
        s = w.sum()
        mu = (w[:, None] * x[t]).sum(axis=0) / s
        d = x[t] - mu
        n = np.count_nonzero(w)
        out.append((d * w[:, None]).T @ d * (n / ((n - 1) * s)))
    return out


def weighted_quantile(x, we, q=None):
    """weighted_quantile(x, we, q) / weighted_quantile(sol, q)  filtering.jl:592-595: StatsBase.quantile with
    ProbabilityWeights per state component and time step -> [T][nx]."""
    if q is None:
        x, we, q = x.x, x.we, we
    x, we = np.asarray(x), np.asarray(we)
    T, N, nx = x.shape
    out = np.zeros((T, nx))
    for t in range(T):
        for i in range(nx):
            out[t, i] = _sb_quantile(x[t, :, i], we[t], float(q))
    return out


def _sb_quantile(v, w, p):
    """StatsBase.quantile(v, w::ProbabilityWeights, p) — third-party arithmetic behind filtering.jl:593 (StatsBase is a
    dependency of the reference, not vendored): drop zero weights, sort, h = p (sum(w) - w_1) + w_1, walk the cumulative
    weights S_k while S_k <= h, interpolate v_{k-1} + (h - S_{k-1}) / (S_k - S_{k-1}) (v_k - v_{k-1})."""
    keep = w != 0
    v, w = v[keep], w[keep]
    order = np.argsort(v, kind="stable")
    v, w = v[order], w[order]
    N = v.size
    h = p * (w.sum() - w[0]) + w[0]
    Sk = Skold = 0.0
    vk = vkold = 0.0
    k = 0
    while Sk <= h:
        k += 1
        if k > N:
            return float(v[-1])
        Skold, vkold = Sk, vk
        vk = v[k - 1]
        Sk += w[k - 1]
    return float(vkold + (h - Skold) / (Sk - Skold) * (vk - vkold))
