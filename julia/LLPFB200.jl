# LLPFB200.jl — the reference-side binding of libllpf_b200.so (ccall).  NOT runnable in the build image
# (no julia there); shipped so that a maintainer with Julia can drop the GPU path behind the package's own
# generic functions.  Every method cites the reference method it shadows.
module LLPFB200

using LowLevelParticleFilters, StaticArrays, LinearAlgebra, Statistics
import LowLevelParticleFilters: reset!, predict!, correct!, update!, forward_trajectory, loglik, particles, weights,
    expweights, num_particles, effective_particles, shouldresample, weighted_mean, index, state

const lib = get(ENV, "LLPF_LIB_PATH", joinpath(@__DIR__, "..", "lowlevelparticlefilters.jl_b200", "csrc", "libllpf_b200.so"))

# ---- include/llpf.h PODs (field order and types must match the header) ---------------------------------
struct LLPFModel
    nx::Int32; nu::Int32; ny::Int32; dynamics::Int32
    A::Ptr{Float64}; B::Ptr{Float64}; C::Ptr{Float64}; R1::Ptr{Float64}; R2::Ptr{Float64}
    mu0::Ptr{Float64}; Sigma0::Ptr{Float64}
    dyn_params::NTuple{8,Float64}
    t_switch::Float64; a1_factor::Float64; integ_Ts::Float64
    supersample::Int32; _pad::Int32
end
struct LLPFConfig
    N::Int64; filter::Int32; resampling::Int32
    resample_threshold::Float64; Ts::Float64; seed::UInt64
    scan_mode::Int32; device::Int32; rank::Int32; world::Int32
    particle_dtype::Int32      # 0 = Float64 particles, 1 = Float32 (nx, ny <= 64, linear-Gaussian)
    single_block::Int32        # 1: one thread block runs the whole filter (PMMH-sized N; required by llpf_run_batch)
    metropolis_steps::Int32    # LLPF_RESAMPLE_METROPOLIS: proposals per slot (0 = 32)
    _reserved::Int32
end
struct LLPFHistStats
    xmean::Ptr{Float64}; xmode::Ptr{Float64}; xcov::Ptr{Float64}; q::Ptr{Float64}; nq::Int32; _pad::Int32; xquantile::Ptr{Float64}
end
struct LLPFRunOutputs
    ll_steps::Ptr{Float64}; ess_steps::Ptr{Float64}; resampled::Ptr{Int32}; xhat::Ptr{Float64}
    x_hist::Ptr{Float64}; w_hist::Ptr{Float64}; we_hist::Ptr{Float64}
end

# ---- model descriptors (closures cannot cross a C-ABI) ---------------------------------------------------
"dynamics(x,u,p,t) = A*x .+ B*u ; measurement = C*x ; df = N(0,R1), dg = N(0,R2), d0 = N(mu0,Sigma0)"
struct LGModel
    A::Matrix{Float64}; B::Matrix{Float64}; C::Matrix{Float64}
    R1::Matrix{Float64}; R2::Matrix{Float64}; mu0::Vector{Float64}; Sigma0::Matrix{Float64}
end
"rk4(quadtank, Ts; supersample) with additive N(0,R1) noise, y = C*x + N(0,R2)  (examples/example_quadtank.jl)"
struct QuadtankModel
    p::NTuple{6,Float64}; Ts::Float64; supersample::Int; t_switch::Float64; a1_factor::Float64
    C::Matrix{Float64}; R1::Matrix{Float64}; R2::Matrix{Float64}; mu0::Vector{Float64}; Sigma0::Matrix{Float64}
end

check(rc) = rc == 0 ? nothing : error("llpf status $rc: " * unsafe_string(ccall((:llpf_last_error, lib), Cstring, ())))

mutable struct GPUParticleFilter{M} <: LowLevelParticleFilters.AbstractParticleFilter
    h::Ptr{Cvoid}
    model::M
    N::Int                           # LOCAL particle count (accessor arrays have this length)
    N_global::Int                    # global particle count (history buffers of the trajectory drivers are indexed globally)
    nx::Int; nu::Int; ny::Int
    Ts::Float64; resample_threshold::Float64
    kind::Int32                      # 0 PF, 1 Advanced, 2 Aux, 3 Aux{Advanced}
    epoch::UInt64
end

function c_model(m::LGModel)
    nx, nu, ny = size(m.A, 1), size(m.B, 2), size(m.C, 1)
    LLPFModel(nx, nu, ny, 0, pointer(m.A), pointer(m.B), pointer(m.C), pointer(m.R1), pointer(m.R2), pointer(m.mu0),
              pointer(m.Sigma0), ntuple(_ -> 0.0, 8), Inf, 1.0, 1.0, 1, 0), nx, nu, ny
end
function c_model(m::QuadtankModel)
    LLPFModel(4, 2, 2, 1, C_NULL, C_NULL, pointer(m.C), pointer(m.R1), pointer(m.R2), pointer(m.mu0), pointer(m.Sigma0),
              (m.p..., 0.0, 0.0), m.t_switch, m.a1_factor, m.Ts, m.supersample, 0), 4, 2, 2
end

"ParticleFilter(N, ...)  src/PFtypes.jl:65-75 (kind=0) / AdvancedParticleFilter :200-210 (kind=1) / AuxiliaryParticleFilter :38-49 (kind=2)"
function GPUParticleFilter(N::Integer, m; kind=0, resample_threshold=(kind == 1 ? 0.5 : 0.1), Ts=1.0, seed=0,
                           resampling=0, scan_mode=0, device=0, rank=0, world=1, particle_eltype=Float64,
                           single_block=false, metropolis_steps=0)
    mdl, nx, nu, ny = c_model(m)
    cfg = LLPFConfig(N, kind, resampling, resample_threshold, Ts, seed, scan_mode, device, rank, world,
                     particle_eltype === Float32 ? 1 : 0, single_block ? 1 : 0, metropolis_steps, 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve m check(ccall((:llpf_create, lib), Cint, (Ref{LLPFConfig}, Ref{LLPFModel}, Ref{Ptr{Cvoid}}), cfg, mdl, h))
    pf = GPUParticleFilter(h[], m, Int(N) ÷ world, Int(N), nx, nu, ny, Float64(Ts), Float64(resample_threshold), Int32(kind), UInt64(0))
    finalizer(p -> ccall((:llpf_destroy, lib), Cint, (Ptr{Cvoid},), p.h), pf)
end

# ---- the reference's constructors, overloaded for descriptor-typed arguments (SURVEY §7.2) -----------------------------
# ParticleFilter(N, dynamics, measurement, df, dg, d0; kw...) src/PFtypes.jl:65-75 requires `measurement::Function`; with the
# descriptor types below the same call builds the GPU filter instead.  Densities: anything with `cov` / `mean`.
struct LinearDynamics; A::Matrix{Float64}; B::Matrix{Float64}; end          # dynamics(x,u,p,t) = A*x .+ B*u
struct LinearMeasurement; C::Matrix{Float64}; end                           # measurement(x,u,p,t) = C*x
_cov(d) = Matrix{Float64}(cov(d)); _mean(d) = Vector{Float64}(mean(d))
function LowLevelParticleFilters.ParticleFilter(N::Integer, dyn::LinearDynamics, meas::LinearMeasurement, df, dg, d0;
                                                resample_threshold=0.1, Ts=1.0, seed=0, kwargs...)
    GPUParticleFilter(N, LGModel(dyn.A, dyn.B, meas.C, _cov(df), _cov(dg), _mean(d0), _cov(d0)); kind=0, resample_threshold, Ts, seed, kwargs...)
end
LowLevelParticleFilters.AuxiliaryParticleFilter(N::Integer, dyn::LinearDynamics, meas::LinearMeasurement, df, dg, d0; kw...) =
    GPUParticleFilter(N, LGModel(dyn.A, dyn.B, meas.C, _cov(df), _cov(dg), _mean(d0), _cov(d0)); kind=2, kw...)

# ---- user-defined models: dynamics / measurement_likelihood as CUDA device source (llpf_create_user) ---------------------
"""
    GPUParticleFilter(N, cuda_source::String, p::Vector{Float64}; nx, nu, ny, R1, mu0, Sigma0, kind=1, kw...)
`cuda_source` defines `llpf_user::dynamics<nx>` and `llpf_user::loglik<nx>` (include/llpf.h): the reference's closures
`dynamics(x,u,p,t)` / `measurement_likelihood(x,u,y,p,t)` (src/PFtypes.jl:232,255), compiled at run time into the sweep.
"""
function GPUParticleFilter(N::Integer, cuda_source::String, p::Vector{Float64}; nx, nu, ny, R1, mu0, Sigma0, kind=1,
                           resample_threshold=0.5, Ts=1.0, seed=0, resampling=0, scan_mode=0, device=0)
    R1m, mu, S0 = Matrix{Float64}(R1), Vector{Float64}(mu0), Matrix{Float64}(Sigma0)
    mdl = LLPFModel(nx, nu, ny, 2, C_NULL, C_NULL, C_NULL, pointer(R1m), C_NULL, pointer(mu), pointer(S0), ntuple(_ -> 0.0, 8),
                    Inf, 1.0, 1.0, 1, 0)
    cfg = LLPFConfig(N, kind, resampling, resample_threshold, Ts, seed, scan_mode, device, 0, 1, 0, 0, 0, 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve R1m mu S0 p check(ccall((:llpf_create_user, lib), Cint,
        (Ref{LLPFConfig}, Ref{LLPFModel}, Cstring, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}), cfg, mdl, cuda_source, p, length(p), h))
    pf = GPUParticleFilter(h[], (cuda_source, p), Int(N), Int(N), nx, nu, ny, Float64(Ts), Float64(resample_threshold), Int32(kind), UInt64(0))
    finalizer(f -> ccall((:llpf_destroy, lib), Cint, (Ptr{Cvoid},), f.h), pf)
end
"per-call parameter override `p` (src/filtering.jl:140,164) of a user-defined model"
set_params!(pf::GPUParticleFilter, p::Vector{Float64}) =
    check(ccall((:llpf_set_user_params, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32), pf.h, p, length(p)))

# ---- verbs -------------------------------------------------------------------------------------------------
"reset!(pf)  src/filtering.jl:4-14"
function reset!(pf::GPUParticleFilter)
    pf.epoch += 1
    check(ccall((:llpf_reset, lib), Cint, (Ptr{Cvoid}, UInt64), pf.h, pf.epoch))
end
"index(pf) = state.t[]  src/PFtypes.jl:319"
function index(pf::GPUParticleFilter)
    t = Ref{Int64}(0); check(ccall((:llpf_index, lib), Cint, (Ptr{Cvoid}, Ref{Int64}), pf.h, t)); Int(t[])
end
vecf(v) = collect(Float64, v)
"correct!(pf,u,y,p,t) -> (ll,0)  src/filtering.jl:164-174"
function correct!(pf::GPUParticleFilter, u, y, p=nothing, t=index(pf) * pf.Ts)
    ll = Ref{Float64}(0.0)
    uv, yv = vecf(u), vecf(y)     # passed as Ptr{Float64}: ccall roots array arguments for the duration of the call
    check(ccall((:llpf_correct, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Float64}), pf.h, uv, yv, t, ll))
    ll[], 0
end
"predict!(pf,u,p,t)  src/filtering.jl:140-153"
predict!(pf::GPUParticleFilter, u, p=nothing, t=index(pf) * pf.Ts) =
    check(ccall((:llpf_predict, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), pf.h, vecf(u), t))
"predict!(pfa,u,y1,p,t)  src/filtering.jl:195-234"
predict!(pf::GPUParticleFilter, u, y1, p, t) =
    check(ccall((:llpf_predict_aux, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), pf.h, vecf(u), vecf(y1), t))
"update!(f,u,y,p,t)  src/filtering.jl:181-185"
function update!(pf::GPUParticleFilter, u, y, p=nothing, t::Real=index(pf) * pf.Ts)
    ll = Ref{Float64}(0.0)
    uv, yv = vecf(u), vecf(y)
    GC.@preserve uv yv check(ccall((:llpf_update, lib), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Float64}), pf.h, uv, yv, C_NULL, t, ll))
    ll[], 0
end
"update!(pfa,u,y,y1,p,t)  src/filtering.jl:187-191 (y1 positional, like the reference)"
function update!(pf::GPUParticleFilter, u, y, y1::AbstractVector, p, t::Real=index(pf) * pf.Ts)
    ll = Ref{Float64}(0.0)
    uv, yv, y1v = vecf(u), vecf(y), vecf(y1)
    GC.@preserve uv yv y1v check(ccall((:llpf_update, lib), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Float64}), pf.h, uv, yv, y1v, t, ll))
    ll[], 0
end
(pf::GPUParticleFilter)(u, y, p=nothing, t=index(pf) * pf.Ts) = update!(pf, u, y, p, t)                 # src/filtering.jl:238
(pf::GPUParticleFilter)(u, y, y1::AbstractVector, p, t=index(pf) * pf.Ts) = update!(pf, u, y, y1, p, t)  # :239

flat(v::AbstractVector) = reduce(hcat, v)          # n x T column-major == the memory of Vector{SVector{n}}

"loglik(pf,u,y,p)  src/smoothing.jl:227-236"
function loglik(pf::GPUParticleFilter, u::AbstractVector, y::AbstractVector, p=nothing)
    U, Y = flat(u), flat(y); ll = Ref{Float64}(0.0); pf.epoch += 1
    check(ccall((:llpf_run, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int32, UInt64, Ref{Float64}, Ptr{Cvoid}),
                pf.h, length(y), U, Y, 1, pf.epoch, ll, C_NULL))
    ll[]
end
"forward_trajectory(pf,u,y,p) -> ParticleFilteringSolution  src/filtering.jl:343-384, src/solutions.jl:334-345"
function forward_trajectory(pf::GPUParticleFilter, u::AbstractVector, y::AbstractVector, p=nothing)
    U, Y = flat(u), flat(y); T = length(y)
    N = pf.N_global      # the device history is indexed by GLOBAL particle number; a sharded rank fills its own rows
    x = Matrix{SVector{pf.nx,Float64}}(undef, N, T); w = Matrix{Float64}(undef, N, T); we = similar(w)
    out = LLPFRunOutputs(C_NULL, C_NULL, C_NULL, C_NULL, pointer(reinterpret(Float64, x)), pointer(w), pointer(we))
    ll = Ref{Float64}(0.0); pf.epoch += 1
    GC.@preserve x w we check(ccall((:llpf_run, lib), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int32, UInt64, Ref{Float64}, Ref{LLPFRunOutputs}),
        pf.h, T, U, Y, 0, pf.epoch, ll, out))
    LowLevelParticleFilters.ParticleFilteringSolution(pf, u, y, x, w, we, ll[])
end

"xb, ll = smooth(pf, M, u, y, p)  src/smoothing.jl:104-143 (forward filtering, backward simulation; history stays on the GPU)"
function LowLevelParticleFilters.smooth(pf::GPUParticleFilter, M::Integer, u::AbstractVector, y::AbstractVector, p=nothing)
    U, Y = flat(u), flat(y); T = length(y)
    xb = Matrix{SVector{pf.nx,Float64}}(undef, M, T)       # M x T, column-major == [T][M][nx] doubles
    ll = Ref{Float64}(0.0); pf.epoch += 1
    GC.@preserve xb check(ccall((:llpf_smooth, lib), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, UInt64, Ref{Float64}, Ptr{Float64}, Ptr{Cvoid}),
        pf.h, T, U, Y, M, pf.epoch, ll, pointer(reinterpret(Float64, xb)), C_NULL))
    xb, ll[]
end

"""
    lls = loglik_batch(pfs::Vector{<:GPUParticleFilter}, u, y)
[loglik(pf, u, y) for pf in pfs] in ONE kernel launch (llpf_run_batch; filters created with single_block=true): what
`metropolis_threaded` (src/smoothing.jl:335-347) needs per MCMC iteration.
"""
function loglik_batch(pfs::Vector{<:GPUParticleFilter}, u::AbstractVector, y::AbstractVector)
    U, Y = flat(u), flat(y); hs = [pf.h for pf in pfs]
    epochs = UInt64[(pf.epoch += 1) for pf in pfs]; lls = zeros(length(pfs))
    GC.@preserve U Y hs epochs lls check(ccall((:llpf_run_batch, lib), Cint,
        (Int32, Ptr{Ptr{Cvoid}}, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{UInt64}, Ptr{Float64}),
        length(pfs), hs, length(y), U, Y, 1, epochs, lls))
    lls
end

"""
    ll, xmean, xmode, xcov, xq = trajectory_statistics(pf, u, y; q=Float64[])
forward_trajectory reduced on the device to mean_trajectory / mode_trajectory / weighted_cov / weighted_quantile
(src/filtering.jl:417-440, 575-595) — the N x T history never leaves HBM (llpf_run_stats).
"""
function trajectory_statistics(pf::GPUParticleFilter, u::AbstractVector, y::AbstractVector; q::Vector{Float64}=Float64[])
    U, Y = flat(u), flat(y); T = length(y); nx = pf.nx
    xmean = zeros(nx, T); xmode = zeros(nx, T); xcov = zeros(nx, nx, T); xq = zeros(nx, length(q), T)
    st = LLPFHistStats(pointer(xmean), pointer(xmode), pointer(xcov), pointer(q), length(q), 0, pointer(xq))
    ll = Ref{Float64}(0.0); pf.epoch += 1
    GC.@preserve U Y xmean xmode xcov xq q check(ccall((:llpf_run_stats, lib), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, UInt64, Ref{Float64}, Ptr{Cvoid}, Ref{LLPFHistStats}),
        pf.h, T, U, Y, pf.epoch, ll, C_NULL, st))
    ll[], xmean, xmode, xcov, xq
end

# ---- accessors (src/PFtypes.jl:296-334, src/resample.jl:1-10, src/filtering.jl:541-568) --------------------
num_particles(pf::GPUParticleFilter) = pf.N
function particles(pf::GPUParticleFilter)
    x = Vector{SVector{pf.nx,Float64}}(undef, pf.N)
    check(ccall((:llpf_get_particles, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), pf.h, pointer(reinterpret(Float64, x)))); x
end
getvec(pf, sym) = (v = Vector{Float64}(undef, pf.N); check(ccall((sym, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), pf.h, v)); v)
weights(pf::GPUParticleFilter) = getvec(pf, :llpf_get_weights)
expweights(pf::GPUParticleFilter) = getvec(pf, :llpf_get_expweights)
function effective_particles(pf::GPUParticleFilter)
    v = Ref{Float64}(0.0); check(ccall((:llpf_effective_particles, lib), Cint, (Ptr{Cvoid}, Ref{Float64}), pf.h, v)); v[]
end
function shouldresample(pf::GPUParticleFilter)
    v = Ref{Int32}(0); check(ccall((:llpf_shouldresample, lib), Cint, (Ptr{Cvoid}, Ref{Int32}), pf.h, v)); v[] != 0
end
function weighted_mean(pf::GPUParticleFilter)
    v = Vector{Float64}(undef, pf.nx); check(ccall((:llpf_weighted_mean, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), pf.h, v)); v
end

"resample(ResampleSystematic, we, j, bins, M) src/resample.jl:17-36 with the rand() of :23 supplied"
function resample_systematic(we::Vector{Float64}, u01::Float64; M=length(we), j=collect(Int64, 1:M), scan_mode=0, device=0)
    bins = similar(we)
    check(ccall((:llpf_resample_systematic, lib), Cint, (Int64, Ptr{Float64}, Float64, Int64, Ptr{Int64}, Ptr{Float64}, Int32, Int32),
                length(we), we, u01, M, j, bins, scan_mode, device))
    j, bins
end

"resample(ResampleStratified, ...) src/resample.jl:38-61 with the M rand() draws of :49 supplied"
function resample_stratified(we::Vector{Float64}, u01::Vector{Float64}; M=length(we), j=collect(Int64, 1:M), scan_mode=0, device=0)
    bins = similar(we)
    check(ccall((:llpf_resample_stratified, lib), Cint, (Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Int32, Int32),
                length(we), we, u01, M, j, bins, scan_mode, device))
    j, bins
end

"resample(ResampleResidual, ...) src/resample.jl:63-117 with the rand() draws of :106 supplied in draw order (length M)"
function resample_residual(we::Vector{Float64}, u01::Vector{Float64}; M=length(we), j=collect(Int64, 1:M), scan_mode=0, device=0)
    bins = similar(we)
    check(ccall((:llpf_resample_residual, lib), Cint, (Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Float64}, Int32, Int32),
                length(we), we, u01, M, j, bins, scan_mode, device))
    j, bins
end

# ---- Rao-Blackwellized particle filter (src/rbpf.jl) on the user-model boundary ------------------------------------------
_lit(v) = string(Float64(v))          # shortest round-trip decimal: parses back to the same Float64
function _cmat(M, rows, cols)
    A = zeros(rows, max(cols, 1))
    (cols > 0 && M !== nothing) && (A[:, 1:cols] .= reshape(Matrix{Float64}(M), rows, cols))
    "{" * join(["{" * join(_lit.(A[r, :]), ", ") * "}" for r in 1:rows], ", ") * "}"
end
"""
    rbpf_source(nxn, nxl, ny, nu, A, B, C, An, R1l, R1n, R2, fn_body, g_body)
The `llpf_user` source of an RBPF (include/llpf.h, LLPF_USER_STATE_HOOKS): a particle is `[xn; xl; tril(R)]`, the arithmetic of
src/rbpf.jl:163-283 is csrc/llpf_rbpf.cuh.  `fn_body` writes fn(xn,u,p,t) into `fn[]`, `g_body` writes g(xn,u,p,t) into `yn[]`.
"""
function rbpf_source(nxn, nxl, ny, nu, A, B, C, An, R1l, R1n, R2, fn_body, g_body)
    nx = nxn + nxl + nxl * (nxl + 1) ÷ 2
    zan = An === nothing || iszero(An); zc = C === nothing || iszero(C)
    k = join([_cmat(A, nxl, nxl), _cmat(B, nxl, nu), _cmat(C, ny, nxl), _cmat(An, nxn, nxl), _cmat(R1l, nxl, nxl),
              _cmat(R1n, nxn, nxn), _cmat(R2, ny, ny), string(Int(zan)), string(Int(zc))], ", ")
    T = "$nxn, $nxl, $ny, $nu, $nx"
    """
    // LLPF_USER_STATE_HOOKS : Rao-Blackwellized particle filter (rbpf.jl)
    #include "llpf_rbpf.cuh"
    namespace llpf_user {
    using RB = llpf_rbpf::Consts<$nxn, $nxl, $ny, $nu>;
    __device__ __forceinline__ RB rb_consts() { const RB k = {$k}; return k; }
    __device__ __forceinline__ void rb_fn(const double (&xn)[$nxn], const double* u, const double* p, double t, double (&fn)[$nxn]) { $fn_body }
    __device__ __forceinline__ void rb_g(const double (&xn)[$nxn], const double* u, const double* p, double t, double (&yn)[$ny]) { $g_body }
    __device__ __forceinline__ void rb_xn(const double (&x)[$nx], double (&xn)[$nxn]) { for (int r = 0; r < $nxn; ++r) xn[r] = x[r]; }
    template <> __device__ void dynamics<$nx>(double (&x)[$nx], const double* u, const double* p, double t) {
      const RB k = rb_consts(); double xn[$nxn], fn[$nxn]; rb_xn(x, xn); rb_fn(xn, u, p, t, fn); llpf_rbpf::predict_mean<$T>(x, fn, u, k); }
    template <> __device__ void add_noise<$nx>(double (&x)[$nx], const double (&xprev)[$nx], const double (&nz)[$nx], const double* u, const double* p, double t) {
      const RB k = rb_consts(); llpf_rbpf::add_noise<$T>(x, xprev, nz, k); }
    template <> __device__ double loglik<$nx>(const double (&x)[$nx], const double* u, const double* y, const double* p, double t) {
      const RB k = rb_consts(); double xn[$nxn], yn[$ny]; rb_xn(x, xn); rb_g(xn, u, p, t, yn); return llpf_rbpf::loglik<$T>(x, yn, y, k); }
    template <> __device__ void correct_state<$nx>(double (&x)[$nx], const double* u, const double* y, const double* p, double t) {
      const RB k = rb_consts(); double xn[$nxn], yn[$ny]; rb_xn(x, xn); rb_g(xn, u, p, t, yn); llpf_rbpf::correct_state<$T>(x, yn, y, k); }
    }
    """
end
"""
    RBPF(N, kf::KalmanFilter, fn_body::String, g_body::String, R1n, d0n; An=nothing, nu, ny, kw...)
RBPF(N, kf, dynamics, nl_measurement_model, R1n, d0n; An, nu) of src/rbpf.jl:113-133 with the two closures given as device
snippets and constant matrices taken from `kf` (A, B, C, R1, R2, d0).  Returns a GPUParticleFilter: every generic verb applies.
"""
function RBPF(N::Integer, kf, fn_body::String, g_body::String, R1n, d0n; An=nothing, nu::Int, ny::Int=size(kf.R2, 1),
              resample_threshold=0.1, Ts=1.0, seed=0, p=Float64[], kw...)
    nxn, nxl = length(_mean(d0n)), length(_mean(kf.d0))
    nx = nxn + nxl + nxl * (nxl + 1) ÷ 2
    nx <= 8 || error("RBPF particle [xn; xl; tril(R)] has $nx components; the f64 engine carries at most 8")
    src = rbpf_source(nxn, nxl, ny, nu, kf.A, nu > 0 ? kf.B : nothing, kf.C, An, kf.R1, R1n, kf.R2, fn_body, g_body)
    R1c = zeros(nx, nx); R1c[1:nxn, 1:nxn] .= Matrix{Float64}(R1n)
    mu0 = zeros(nx); S0 = zeros(nx, nx)
    mu0[1:nxn] .= _mean(d0n); S0[1:nxn, 1:nxn] .= _cov(d0n); mu0[nxn+1:nxn+nxl] .= _mean(kf.d0)
    R0 = Matrix{Float64}(_cov(kf.d0))
    mu0[nxn+nxl+1:end] .= [R0[r, c] for r in 1:nxl for c in 1:r]
    GPUParticleFilter(N, src, Vector{Float64}(p); nx, nu, ny, R1=R1c, mu0, Sigma0=S0, kind=0, resample_threshold, Ts, seed, kw...)
end

# ---- Ensemble Kalman filter (src/enkf.jl): the llpf_enkf_* verbs on an ordinary handle --------------------------------------
mutable struct GPUEnsembleKalmanFilter <: LowLevelParticleFilters.AbstractKalmanFilter
    pf::GPUParticleFilter          # the handle: its particle buffer is the ensemble
    inflation::Float64
end
"EnsembleKalmanFilter(dynamics, measurement, R1, R2, d0, N; nu, ny, inflation)  src/enkf.jl:94-141 (descriptor arguments)"
function LowLevelParticleFilters.EnsembleKalmanFilter(dyn::LinearDynamics, meas::LinearMeasurement, R1, R2, d0, N::Integer;
                                                      inflation=1.0, Ts=1.0, seed=0, kw...)
    pf = GPUParticleFilter(N, LGModel(dyn.A, dyn.B, meas.C, Matrix{Float64}(R1), Matrix{Float64}(R2), _mean(d0), _cov(d0)); kind=0, Ts, seed)
    check(ccall((:llpf_enkf_set_inflation, lib), Cint, (Ptr{Cvoid}, Float64), pf.h, inflation))
    check(ccall((:llpf_enkf_reset, lib), Cint, (Ptr{Cvoid}, UInt64), pf.h, 0))
    GPUEnsembleKalmanFilter(pf, inflation)
end
function _enkf_state(e::GPUEnsembleKalmanFilter)
    m = zeros(e.pf.nx); c = zeros(e.pf.nx, e.pf.nx); t = Ref{Int64}(0)
    check(ccall((:llpf_enkf_state, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}), e.pf.h, m, c, t))
    m, c, Int(t[])
end
LowLevelParticleFilters.state(e::GPUEnsembleKalmanFilter) = _enkf_state(e)[1]            # enkf.jl:186
LowLevelParticleFilters.covariance(e::GPUEnsembleKalmanFilter) = _enkf_state(e)[2]       # enkf.jl:193 (symmetric)
index(e::GPUEnsembleKalmanFilter) = _enkf_state(e)[3]
particles(e::GPUEnsembleKalmanFilter) = particles(e.pf)
num_particles(e::GPUEnsembleKalmanFilter) = e.pf.N
function reset!(e::GPUEnsembleKalmanFilter)                                              # enkf.jl:205-224
    e.pf.epoch += 1
    check(ccall((:llpf_enkf_reset, lib), Cint, (Ptr{Cvoid}, UInt64), e.pf.h, e.pf.epoch))
end
predict!(e::GPUEnsembleKalmanFilter, u, p=nothing, t::Real=index(e) * e.pf.Ts) =         # enkf.jl:228-272
    check(ccall((:llpf_enkf_predict, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), e.pf.h, vecf(u), t))
function correct!(e::GPUEnsembleKalmanFilter, u, y, p=nothing, t::Real=index(e) * e.pf.Ts)   # enkf.jl:281-356
    nx, ny = e.pf.nx, e.pf.ny
    ll = Ref{Float64}(0.0); ev = zeros(ny); S = zeros(ny, ny); K = zeros(ny, nx)        # K arrives row-major nx x ny
    check(ccall((:llpf_enkf_correct, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ref{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}), e.pf.h, vecf(u), vecf(y), t, ll, ev, S, K))
    (; ll=ll[], e=ev, S, Sᵪ=cholesky(Symmetric(S)), K=permutedims(K))
end
function update!(e::GPUEnsembleKalmanFilter, u, y, p=nothing, t::Real=index(e) * e.pf.Ts)   # enkf.jl:361-366
    r = correct!(e, u, y, p, t); predict!(e, u, p, t); r
end
function forward_trajectory(e::GPUEnsembleKalmanFilter, u::AbstractVector, y::AbstractVector, p=nothing)   # filtering.jl:282-325
    T, nx, ny = length(y), e.pf.nx, e.pf.ny
    U = e.pf.nu > 0 ? Matrix{Float64}(reduce(hcat, u)) : zeros(0, T); Y = Matrix{Float64}(reduce(hcat, y))
    x = zeros(nx, T); xt = zeros(nx, T); R = zeros(nx, nx, T); Rt = zeros(nx, nx, T); ev = zeros(ny, T); lls = zeros(T)
    S = zeros(ny, ny, T); K = zeros(ny, nx, T); ll = Ref{Float64}(0.0)
    e.pf.epoch += 1
    check(ccall((:llpf_enkf_run, lib), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, UInt64, Ref{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                e.pf.h, T, U, Y, e.pf.epoch, ll, x, R, xt, Rt, ev, lls, S, K))
    cols(M) = [M[:, k] for k in 1:T]; mats(M) = [M[:, :, k] for k in 1:T]
    LowLevelParticleFilters.KalmanFilteringSolution(e, u, y, cols(x), cols(xt), mats(R), mats(Rt), ll[], cols(ev),
        [permutedims(K[:, :, k]) for k in 1:T], [cholesky(Symmetric(S[:, :, k])) for k in 1:T], nothing, range(0, step=e.pf.Ts, length=T))
end

end # module
