// llpf_engine_batch.cu — the batched multi-chain engine: k_engine_batch<NX, NY, DYN, RESID> (llpf_engine.cuh, -DLLPF_BATCH:
// one thread block = one independent single-block filter).  The PMMH driver of the reference (src/smoothing.jl:266-347,
// examples/example_lineargaussian.jl:195-223) evaluates `loglik` thousands of times on filters of ~1000 particles; one
// such filter occupies one SM-half for ~1 ms, so hundreds of chains are evaluated by ONE launch (llpf_run_batch).
#define LLPF_BATCH 1
#include "llpf_engine.cuh"

namespace llpf {

#define LLPF_BATCH_LIST(X) X(1, 1, 0, 0) X(2, 1, 0, 0) X(2, 2, 0, 0) X(4, 2, 0, 0)

const void* engine_batch_kernel(int nx, int ny, int dyn, int resid) {
#define X(NX, NY, DYN, R) \
  if (nx == NX && ny == NY && dyn == DYN && resid == R) return (const void*)k_engine_batch<NX, NY, DYN, R>;
  LLPF_BATCH_LIST(X)
#undef X
  return nullptr;
}

}  // namespace llpf
