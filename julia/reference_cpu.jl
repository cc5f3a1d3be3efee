# reference_cpu.jl — times the REAL LowLevelParticleFilters.jl on BASELINE configs 1-3 (needs Julia; not runnable in
# the build image).  Paste the output next to bench.py's cpu_baseline.
using LowLevelParticleFilters, LinearAlgebra, StaticArrays, Distributions, Random
println("threads = ", Threads.nthreads()); versioninfo()
function lg(nx, nu, ny)
    Tr = randn(nx, nx)
    A = SMatrix{nx,nx}(Tr * diagm(0 => LinRange(0.5, 0.95, nx)) / Tr); B = @SMatrix randn(nx, nu); C = @SMatrix randn(ny, nx)
    (; A, B, C, df=MvNormal(Diagonal(ones(nx))), dg=MvNormal(Diagonal(ones(ny))), d0=MvNormal(randn(nx), 4.0 * I))
end
Random.seed!(0)
for (nx, N, T) in ((2, 500, 200), (4, 500, 200), (4, 2^16, 100), (4, 2^20, 20))
    m = lg(nx, 2, 2)
    dyn(x, u, p, t) = m.A * x .+ m.B * u; meas(x, u, p, t) = m.C * x
    pf = ParticleFilter(N, dyn, meas, m.df, m.dg, m.d0)
    x, u, y = LowLevelParticleFilters.simulate(pf, T, MvNormal(Diagonal(ones(2))))
    loglik(pf, u, y)
    t = @elapsed ll = loglik(pf, u, y)
    println("ParticleFilter nx=$nx N=$N T=$T: $(round(N * T / t / 1e6, digits=2)) M particle-steps/s  ll=$ll")
end
