// llpf_metropolis.cuh — Metropolis resampling (Murray, Lee & Jacob, "Parallel resampling in the particle filter", JCGS 2016),
// the low-synchronisation alternative named by the north star.  NOT part of the reference (its strategies are
// Systematic / Stratified / Residual, src/LowLevelParticleFilters.jl:43-46): an extension, tested statistically only.
//
//   for every output slot i:  k = i ;  repeat B times:  j ~ U{1..N}, u ~ U(0,1) ;  if u <= w_j / w_k : k = j ;  ancestor(i) = k
//
// Only weight RATIOS are used: no prefix sum, no normalisation, no ordering — one grid barrier (the weights are stashed
// so that the fused sweep may overwrite w while other blocks still walk their chains) instead of the scan's two barriers
// and its cross-GPU exchanges.  The price: B dependent random reads per slot and a bias that decays geometrically in B
// (B = llpf_config.metropolis_steps, default 32).  Single-GPU filters.  RNG: stream 7, counter (step, slot, proposal pair).
#pragma once

namespace llpf {

constexpr uint32_t ST_METRO = 7;

// lw: stash of the raw log-weights (any common offset cancels), n = N particles; returns the ancestor (0-based) of slot gi
__device__ __forceinline__ int metropolis_chain(const double* __restrict__ lw, int n, int gi, int B, const RngKey& key,
                                                uint32_t step_idx, const MathTab& T) {
  int k = gi;
  double wk = __ldcg(lw + k);
  for (int b = 0; 2 * b < B; ++b) {
    const uint4 r = rng_block(key, ST_METRO, step_idx, (unsigned long long)(unsigned)gi, (uint32_t)b);
    const uint32_t ru[2] = {r.y, r.w};
    double lu[2];
    log_u32_v<2>(ru, lu, T);                       // ln u, u = (r + 1/2) 2^-32
    const int j0 = (int)__umulhi(r.x, (uint32_t)n), j1 = (int)__umulhi(r.z, (uint32_t)n);
    const double w0 = __ldcg(lw + j0), w1 = __ldcg(lw + j1);   // both proposals' weights in flight together
    if (lu[0] <= w0 - wk) { k = j0; wk = w0; }
    if (2 * b + 1 < B && lu[1] <= w1 - wk) { k = j1; wk = w1; }
  }
  return k;
}

// The whole resample for the block's own slots [beg, end): stash -> grid barrier -> chains -> j (global ids, non-monotone).
// wsrc: raw log-weights (nullptr: all equal).  hist: optional history rows of the normalised weights (already offset).
template <class NormFn>
__device__ __forceinline__ void resample_metropolis(const EngineP& P, Shared& sh, int beg, int end, unsigned& bar_target,
                                                    const double* wsrc, uint32_t step_idx, NormFn histfn) {
  for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
    const double wr = wsrc ? __ldcg(wsrc + i) : 0.0;
    histfn(i, wr);
    __stcg(P.bins + i, wr);
  }
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  const int B = P.metro_steps > 0 ? P.metro_steps : 32;
  for (int i = beg + threadIdx.x; i < end; i += BLOCK)     // same thread <-> slot mapping as the sweep that follows
    __stcg(P.j + i, metropolis_chain(P.bins, P.n, i, B, P.key, step_idx, sh.mt));
}

}  // namespace llpf
