# dump_golden.jl — writes tests/golden/reference_v1.json from the REAL LowLevelParticleFilters.jl (needs Julia + the package;
# there is no julia in the build image, so this file has not been executed there).
#
#     julia --project=<env with LowLevelParticleFilters, Distributions, StaticArrays> julia/dump_golden.jl [out.json]
#
# Purpose: pin the repo's oracle against outputs of the reference itself.  The reference draws randn from `pf.rng`
# (a sequential Xoshiro) and rand() from the task-local global RNG inside `resample` (src/resample.jl:23,49,106); neither can
# be replayed by a counter-based parallel generator.  So this script RECORDS every random variate the reference consumes
#   * initial particles         rand(rng, initial_density)                       src/filtering.jl:8
#   * dynamics noise vectors    rand!(pf.rng, dynamics_density, noise)           src/PFtypes.jl:135,153 ; ext/...DistributionsExt.jl:90
#   * the resampling uniforms   rand()                                           src/resample.jl:23,49,106
#     (recovered by copying the task-local RNG state right before predict! and replaying it — shouldresample draws nothing)
# together with everything the reference computes from them (x, w, we of every step, ll, j) — for the particle filters and,
# in the `rbpf` / `enkf` sections, for the Rao-Blackwellized and the ensemble Kalman filter (consumers: oracle/rbpf_ref.py,
# oracle/enkf_ref.py).  tests/test_reference_golden.py
# feeds the recorded variates to the oracle restatement (oracle/pyref.py, `inject=`) and demands the same outputs:
# indices and resample decisions exactly, floating-point values to 1e-12 relative (SLEEFPirates.exp / Distributions.logpdf
# differ from libm / the Cholesky form in the last bits).
using LowLevelParticleFilters, Distributions, StaticArrays, LinearAlgebra, Random, Statistics
const LLPF = LowLevelParticleFilters

# ---- recording wrappers -------------------------------------------------------------------------------------------------
struct RecordingNoise{D} <: Distributions.Sampleable{Multivariate,Continuous}
    d::D
    log::Vector{Vector{Float64}}
end
Base.length(d::RecordingNoise) = length(d.d)
Base.eltype(d::RecordingNoise) = Float64
function Random.rand!(rng::Random.AbstractRNG, d::RecordingNoise, out::AbstractVector)
    rand!(rng, d.d, out)
    push!(d.log, collect(Float64, out))
    out
end
struct RecordingInit{D} <: Distributions.Sampleable{Multivariate,Continuous}
    d::D
    log::Vector{Vector{Float64}}
end
Base.length(d::RecordingInit) = length(d.d)
Base.eltype(d::RecordingInit) = Float64
function Base.rand(rng::Random.AbstractRNG, d::RecordingInit)
    x = rand(rng, d.d)
    push!(d.log, collect(Float64, x))
    x
end
Base.rand(d::RecordingInit) = rand(Random.default_rng(), d)
Statistics.mean(d::RecordingInit) = mean(d.d)

# ---- tiny JSON writer (round-trip exact: repr(::Float64) is the shortest string that parses back to the same bits) -------
js(x::AbstractFloat) = isfinite(x) ? repr(Float64(x)) : (isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity"))
js(x::Integer) = string(x)
js(x::Bool) = x ? "true" : "false"
js(x::AbstractString) = "\"" * x * "\""
js(x::AbstractVector) = "[" * join((js(v) for v in x), ",") * "]"
js(x::AbstractMatrix) = js([collect(x[i, :]) for i in 1:size(x, 1)])       # list of rows
js(x::SVector) = js(collect(x))
js(d::AbstractDict) = "{" * join(("\"$(k)\":" * js(v) for (k, v) in d), ",") * "}"

# ---- one recorded run ---------------------------------------------------------------------------------------------------
# Drives the filter exactly like forward_trajectory (src/filtering.jl:343-384) / loglik (src/smoothing.jl:227-236), calling
# the reference's own verbs, and records the variates and the state after every correct!.
function replay_uniforms(nmax)
    r = copy(Random.default_rng())      # Xoshiro copy of the task-local RNG state
    [rand(r) for _ in 1:nmax]
end

function recorded_run(pf, dn, d0, u, y; mode::Symbol)
    isaux = pf isa AuxiliaryParticleFilter
    inner = isaux ? pf.pf : pf
    N = num_particles(pf); T = length(y)
    empty!(d0.log); empty!(dn.log)
    reset!(pf)
    x0 = copy(d0.log)
    xs = Vector{Vector{Vector{Float64}}}(); ws = Vector{Vector{Float64}}(); wes = Vector{Vector{Float64}}()
    lls = Float64[]; ures = Vector{Vector{Float64}}(); noise = Vector{Vector{Vector{Float64}}}(); resampled = Int[]
    strategy = LLPF.resampling_strategy(pf)
    ndraw = strategy === ResampleSystematic ? 1 : N
    for t in 1:T
        ti = mode === :forward_trajectory ? (t - 1) * inner.Ts : (isaux ? (t - 1) * inner.Ts : LLPF.index(pf) * inner.Ts)
        last_aux_loglik = isaux && mode === :loglik && t == T
        f = last_aux_loglik ? inner : pf                      # smoothing.jl:235: the INNER filter's update! on the last sample
        lli = correct!(f, u[t], y[t], LLPF.parameters(f), ti)[1]
        push!(lls, lli)
        push!(xs, [collect(Float64, p) for p in LLPF.particles(pf)])
        push!(ws, copy(LLPF.weights(pf))); push!(wes, copy(LLPF.expweights(pf)))
        empty!(dn.log)
        draws = replay_uniforms(ndraw)
        if isaux && !last_aux_loglik
            if t < T
                predict!(pf, u[t], y[t+1], LLPF.parameters(pf), ti)          # always resamples (filtering.jl:205)
                push!(ures, draws); push!(resampled, 1)
            else
                push!(ures, Float64[]); push!(resampled, 0)                  # forward_trajectory(pfa): no predict! at t == T
            end
        else
            did = LLPF.shouldresample(f)
            predict!(f, u[t], LLPF.parameters(f), ti)
            push!(ures, did ? draws : Float64[]); push!(resampled, did ? 1 : 0)
        end
        push!(noise, copy(dn.log))
    end
    Dict("x0" => x0, "noise" => noise, "u_res" => ures, "x" => xs, "w" => ws, "we" => wes, "ll_steps" => lls,
         "ll" => sum(lls), "resampled" => resampled, "j_final" => copy(LLPF.state(pf).j),
         "x_final" => [collect(Float64, p) for p in LLPF.particles(pf)], "w_final" => copy(LLPF.weights(pf)))
end

function lg_case(name, nx, nu, ny, N, T; aux=false, strategy=ResampleSystematic, threshold=0.1, seed=1)
    Random.seed!(seed)
    Tr = randn(nx, nx)
    A = SMatrix{nx,nx}(Tr * diagm(0 => collect(LinRange(0.5, 0.95, nx))) / Tr)
    B = SMatrix{nx,nu}(randn(nx, nu)); C = SMatrix{ny,nx}(randn(ny, nx))
    R1 = Matrix(1.0I, nx, nx); R2 = Matrix(1.0I, ny, ny); mu0 = randn(nx); S0 = Matrix(4.0I, nx, nx)
    dyn(x, u, p, t) = A * x .+ B * u
    meas(x, u, p, t) = C * x
    dn = RecordingNoise(MvNormal(zeros(nx), R1), Vector{Float64}[])
    d0 = RecordingInit(MvNormal(mu0, S0), Vector{Float64}[])
    dg = MvNormal(zeros(ny), R2)
    pf = ParticleFilter(N, dyn, meas, dn, dg, d0; resample_threshold=threshold, resampling_strategy=strategy, ny=ny, nu=nu)
    aux && (pf = AuxiliaryParticleFilter(pf))
    # data: simulate(f,u,p) semantics (src/filtering.jl:462-477), written out so no recording wrapper is consumed
    u = [randn(nu) for _ in 1:T]
    xt = copy(mu0); y = Vector{Vector{Float64}}()
    for t in 1:T
        push!(y, collect(C * xt) .+ randn(ny))
        xt = collect(A * xt .+ B * u[t]) .+ randn(nx)
    end
    us = [SVector{nu}(v) for v in u]; ys = [SVector{ny}(v) for v in y]
    ft = recorded_run(pf, dn, d0, us, ys; mode=:forward_trajectory)
    lk = recorded_run(pf, dn, d0, us, ys; mode=:loglik)
    # the same through the reference's own drivers (consumes fresh variates: only shapes / finiteness are comparable)
    sol = forward_trajectory(pf, us, ys)
    Dict("name" => name, "filter" => aux ? "apf" : "pf",
         "resampling" => strategy === ResampleSystematic ? "systematic" : strategy === ResampleStratified ? "stratified" : "residual",
         "threshold" => threshold, "N" => N, "T" => T, "nx" => nx, "nu" => nu, "ny" => ny, "Ts" => 1.0,
         "A" => Matrix(A), "B" => Matrix(B), "C" => Matrix(C), "R1" => R1, "R2" => R2, "mu0" => mu0, "Sigma0" => S0,
         "u" => u, "y" => y, "forward_trajectory" => ft, "loglik" => lk, "driver_ll_is_finite" => isfinite(sol.ll))
end

function range_vectors()
    # the Float64 range of resample.jl:24 at a few (r, M, total): every element, so that oracle/julia_range.py is pinned
    out = Dict{String,Any}[]
    for (r, M, total) in ((0.05, 10, 1.0), (0.0, 10, 1.0), (0.0, 5, 1.0), (0.25 / 777, 777, 1.0), (0.5 / 1234, 1234, 1.0),
                          (0.05, 10, 0.9999999999999999), (0.0007316351, 1000, 1.0000000000000002), (rand() / 300, 300, 1.0))
        s = r:(1/M):(total+r)
        push!(out, Dict("r" => r, "M" => M, "total" => total, "len" => length(s), "s" => [s[i] for i in 1:M],
                        "ref_hi" => s.ref.hi, "ref_lo" => s.ref.lo, "step_hi" => s.step.hi, "step_lo" => s.step.lo,
                        "offset" => s.offset))
    end
    out
end

function resample_vectors()
    # resample at the function boundary: we, the rand() it consumed, and j (bins is internal: recomputed by the consumer)
    out = Dict{String,Any}[]
    for (N, M, strat) in ((10, 10, ResampleSystematic), (257, 257, ResampleSystematic), (100, 37, ResampleSystematic),
                          (64, 200, ResampleSystematic), (257, 257, ResampleStratified), (100, 100, ResampleResidual))
        w = randn(N) .* 2; we = similar(w); logsumexp!(w, we)
        draws = replay_uniforms(max(N, M))
        j = LLPF.resample(strat, we, fill(-7, M), zeros(N), M)
        push!(out, Dict("N" => N, "M" => M, "strategy" => string(strat), "we" => we, "draws" => draws, "j" => j))
    end
    out
end

# ---- RBPF (src/rbpf.jl): the mixed linear / nonlinear model of test/test_rbpf.jl:5-34 (An != 0) ---------------------------
# rbpf.jl reads the fields μ and Σ of the nonlinear-state distributions (:217, :292-295) and draws rand(pf.rng, d) from them
# (:136-150 reset!, :203 / :222 predict!): a distribution with those fields that logs what it hands out records every variate.
struct RecordingMvN{M,S}
    μ::M
    Σ::S
    log::Vector{Vector{Float64}}
end
Base.length(d::RecordingMvN) = length(d.μ)
Base.eltype(d::RecordingMvN) = Float64
function Base.rand(rng::Random.AbstractRNG, d::RecordingMvN)
    x = d.μ + cholesky(Symmetric(Matrix(d.Σ))).L * randn(rng, length(d.μ))
    push!(d.log, collect(Float64, x))
    typeof(d.μ)(x)
end
flat(p) = vcat(collect(Float64, p.xn), collect(Float64, p.xl), [Float64(p.R[r, c]) for r in 1:size(p.R, 1) for c in 1:r])

function rbpf_case(; N=150, T=40, threshold=0.5, seed=7)
    Random.seed!(seed)
    An = SA[0.5;;]; A = SA[0.95;;]; C2 = SA[1.0;;]; B = @SMatrix zeros(1, 0)
    R1n = SA[0.01;;]; R1l = SA[0.01;;]; R2 = SA[0.1;;]
    x0n = SA[1.0]; x0l = SA[1.0]; R0 = SA[1.0;;]
    fn(xn, args...) = xn
    h(xn, args...) = xn
    dn = RecordingMvN(SA[0.0], R1n, Vector{Float64}[])
    d0n = RecordingMvN(x0n, R1n, Vector{Float64}[])
    kf = KalmanFilter(A, B, C2, 0, R1l, R2, LLPF.SimpleMvNormal(x0l, R0))
    mm = RBMeasurementModel(h, R2, 1)
    pf = RBPF(N, kf, fn, mm, dn, d0n; nu=0, An=An, Ts=1.0, resample_threshold=threshold,
              names=SignalNames(x=["x1", "x2"], u=String[], y=["y1"], name="RBPF"))
    # data from the model itself, written out (no recording wrapper consumed)
    xn, xl = 1.0, 1.0; y = Vector{Vector{Float64}}()
    for t in 1:T
        push!(y, [xn + xl + sqrt(0.1) * randn()])
        xn, xl = xn + 0.5 * xl + 0.1 * randn(), 0.95 * xl + 0.1 * randn()
    end
    us = [SA_F64[] for _ in 1:T]; ys = [SVector{1}(v) for v in y]
    empty!(d0n.log); empty!(dn.log)
    reset!(pf)
    x0 = copy(d0n.log)
    xs = Vector{Vector{Vector{Float64}}}(); ws = Vector{Vector{Float64}}(); wes = Vector{Vector{Float64}}()
    lls = Float64[]; ures = Vector{Vector{Float64}}(); noise = Vector{Vector{Vector{Float64}}}(); resampled = Int[]
    for t in 1:T
        ti = (t - 1) * pf.Ts                                       # forward_trajectory  src/filtering.jl:352
        push!(lls, correct!(pf, us[t], ys[t], LLPF.parameters(pf), ti)[1])
        push!(xs, [flat(p) for p in LLPF.particles(pf)])
        push!(ws, copy(LLPF.weights(pf))); push!(wes, copy(LLPF.expweights(pf)))
        empty!(dn.log)
        draws = replay_uniforms(1)
        did = LLPF.shouldresample(pf)
        predict!(pf, us[t], LLPF.parameters(pf), ti)
        push!(ures, did ? draws : Float64[]); push!(resampled, did ? 1 : 0)
        push!(noise, copy(dn.log))
    end
    Dict("name" => "rbpf_mixed", "N" => N, "T" => T, "threshold" => threshold, "Ts" => 1.0,
         "A" => Matrix(A), "An" => Matrix(An), "C" => Matrix(C2), "R1l" => Matrix(R1l), "R1n" => Matrix(R1n), "R2" => Matrix(R2),
         "x0n" => collect(x0n), "x0l" => collect(x0l), "R0" => Matrix(R0), "y" => y,
         "x0" => x0, "noise" => noise, "u_res" => ures, "x" => xs, "w" => ws, "we" => wes, "ll_steps" => lls, "ll" => sum(lls),
         "resampled" => resampled, "j_final" => copy(LLPF.state(pf).j), "x_final" => [flat(p) for p in LLPF.particles(pf)])
end

# ---- EnKF (src/enkf.jl): the linear test system of test/test_enkf.jl:13-31 -------------------------------------------------
# enkf.jl builds its noise distributions inside predict! / correct! (:245, :337), so wrappers cannot see the draws; an RNG that
# logs every standard normal it hands out can: SimpleMvNormal's rand is mu + cholesky(Sigma).L * randn(rng, n) (src/utils.jl:260).
mutable struct RecordingRNG <: Random.AbstractRNG
    r::Random.Xoshiro
    log::Vector{Float64}
end
Random.randn(g::RecordingRNG) = (v = randn(g.r); push!(g.log, v); v)
Random.randn(g::RecordingRNG, ::Type{Float64}) = randn(g)
Random.rand(g::RecordingRNG, ::Type{T}) where {T} = rand(g.r, T)           # anything else passes through unrecorded
rows(v, n) = [v[(i-1)*n+1:i*n] for i in 1:(length(v) ÷ n)]

function enkf_case(; N=120, T=30, inflation=1.0, seed=11)
    Random.seed!(seed)
    nx, nu, ny = 2, 2, 2
    A = SA[0.99 0.1; 0.0 0.2]; B = SA[-0.74 1.61; -1.44 1.75]; C = SMatrix{2,2}(1.0I(2))
    dynamics(x, u, p, t) = A * x .+ B * u
    measurement(x, u, p, t) = C * x
    R1 = SMatrix{2,2}(1.0I(2)); R2 = SMatrix{2,2}(1.0I(2))
    mu0 = randn(nx); S0 = Matrix(4.0I, nx, nx)
    d0 = LLPF.SimpleMvNormal(SVector{2}(mu0), SMatrix{2,2}(S0))
    g = RecordingRNG(Random.Xoshiro(seed), Float64[])
    enkf = EnsembleKalmanFilter(dynamics, measurement, R1, R2, d0, N; nu, ny, inflation, rng=g)
    u = [randn(nu) for _ in 1:T]
    xt = mu0 .+ 2 .* randn(nx); y = Vector{Vector{Float64}}()
    for t in 1:T
        push!(y, collect(C * xt) .+ randn(ny))
        xt = collect(A * xt .+ B * u[t]) .+ randn(nx)
    end
    us = [SVector{nu}(v) for v in u]; ys = [SVector{ny}(v) for v in y]
    empty!(g.log); reset!(enkf); z0 = rows(copy(g.log), nx)
    xs = Vector{Vector{Float64}}(); Rs = Vector{Vector{Vector{Float64}}}(); xts = Vector{Vector{Float64}}()
    Rts = Vector{Vector{Vector{Float64}}}(); es = Vector{Vector{Float64}}(); lls = Float64[]
    Ss = Vector{Vector{Vector{Float64}}}(); Ks = Vector{Vector{Vector{Float64}}}()
    zobs = Vector{Vector{Vector{Float64}}}(); zdyn = Vector{Vector{Vector{Float64}}}()
    mrows(M) = [collect(Float64, M[i, :]) for i in 1:size(M, 1)]
    for k in 1:T                                                   # forward_trajectory(kf, u, y)  src/filtering.jl:292-318
        ti = (k - 1) * enkf.Ts
        push!(xs, collect(Float64, LLPF.state(enkf))); push!(Rs, mrows(LLPF.covariance(enkf)))
        empty!(g.log)
        ret = correct!(enkf, us[k], ys[k], LLPF.parameters(enkf), ti)
        push!(zobs, rows(copy(g.log), ny))
        push!(lls, ret.ll); push!(es, collect(Float64, ret.e)); push!(Ss, mrows(ret.S)); push!(Ks, mrows(ret.K))
        push!(xts, collect(Float64, LLPF.state(enkf))); push!(Rts, mrows(LLPF.covariance(enkf)))
        empty!(g.log)
        predict!(enkf, us[k], LLPF.parameters(enkf), ti)
        push!(zdyn, rows(copy(g.log), nx))
    end
    Dict("name" => "enkf_lin2", "N" => N, "T" => T, "Ts" => 1.0, "inflation" => inflation, "A" => Matrix(A), "B" => Matrix(B),
         "C" => Matrix(C), "R1" => Matrix(R1), "R2" => Matrix(R2), "mu0" => mu0, "Sigma0" => S0, "u" => u, "y" => y,
         "z0" => z0, "zobs" => zobs, "zdyn" => zdyn, "x" => xs, "R" => Rs, "xt" => xts, "Rt" => Rts, "e" => es, "S" => Ss, "K" => Ks,
         "ll_steps" => lls, "ll" => sum(lls), "ensemble_final" => [collect(Float64, p) for p in LLPF.particles(enkf)])
end

function main()
    out = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "reference_v1.json")
    cases = [
        lg_case("pf_sys_lg4", 4, 2, 2, 200, 30; threshold=0.5, seed=1),
        lg_case("pf_sys_lg2_thr01", 2, 1, 1, 100, 40; threshold=0.1, seed=2),
        lg_case("pf_strat_lg3", 3, 2, 2, 128, 25; strategy=ResampleStratified, threshold=0.5, seed=3),
        lg_case("pf_resid_lg3", 3, 2, 2, 128, 25; strategy=ResampleResidual, threshold=0.5, seed=4),
        lg_case("apf_sys_lg4", 4, 2, 2, 150, 20; aux=true, seed=5),
    ]
    doc = Dict("version" => 1, "julia" => string(VERSION), "package" => (isdefined(Base, :pkgversion) ? string(pkgversion(LLPF)) : "unknown"),
               "cases" => cases, "ranges" => range_vectors(), "resample" => resample_vectors(), "rbpf" => [rbpf_case()],
               "enkf" => [enkf_case(), enkf_case(inflation=1.05, seed=12)])
    open(out, "w") do io
        write(io, js(doc))
    end
    println("wrote ", out)
end
main()
