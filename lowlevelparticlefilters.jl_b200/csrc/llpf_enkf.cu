// llpf_enkf.cu — instantiations of the Ensemble Kalman filter kernel (llpf_enkf.cuh) and their lookup for the C-ABI
// (llpf_enkf_* in llpf_api.cu).  A translation unit of its own: the headline engine kernels are not recompiled with it.
#include "llpf_enkf.cuh"

namespace llpf {

#define LLPF_ENKF_LIST X(1, 1, 0) X(2, 1, 0) X(2, 2, 0) X(3, 1, 0) X(3, 2, 0) X(4, 1, 0) X(4, 2, 0) X(4, 4, 0) X(6, 2, 0) X(8, 2, 0) X(4, 2, 1)

const void* enkf_kernel(int nx, int ny, int dyn) {
#define X(NX, NY, DYN) \
  if (nx == NX && ny == NY && dyn == DYN) return (const void*)k_enkf<NX, NY, DYN>;
  LLPF_ENKF_LIST
#undef X
  return nullptr;
}

}  // namespace llpf
