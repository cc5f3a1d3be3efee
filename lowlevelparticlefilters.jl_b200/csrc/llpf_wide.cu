// llpf_wide.cu — the Float32-particle ("wide") engine as its own translation unit: k_engine_wide is compiled with 128-thread
// blocks, two per SM (__graft_entry__.py passes -DLLPF_BLOCK=128 -DLLPF_MIN_BLOCKS=2), while the rest of the library keeps
// 256.  Two independent blocks per SM run the tile phases out of step — one block's FP64 Box-Muller overlaps the other's
// FFMA2 GEMM — which one 256-thread block with block-wide barriers between the phases cannot do.
#define LLPF_WIDE_ENGINE_TU 1
#include "llpf_wide.cuh"

namespace llpf {
const void* wide_engine_kernel() { return (const void*)k_engine_wide; }
size_t wide_engine_smem_bytes() { return sizeof(WideShared); }
int wide_engine_block_threads() { return BLOCK; }
int wide_engine_min_blocks() { return LLPF_MIN_BLOCKS; }
}  // namespace llpf
