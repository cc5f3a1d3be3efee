// llpf_smooth.cu — instantiations of the FFBS smoother kernel (llpf_smooth.cuh) and their launch entry.
#include <cstring>

#include "llpf_smooth.cuh"

namespace llpf {

template <int NX, int DYN>
static cudaError_t launch_one(const SmoothP& S, const SmoothModelG& G, int grid, cudaStream_t stream) {
  ModelP<NX, 1> M;
  std::memset(&M, 0, sizeof(M));
  for (int r = 0; r < NX; ++r)
    for (int c = 0; c < NX; ++c) {
      M.A[r * NX + c] = G.A[r * MAX_NX + c];
      M.L1[r * NX + c] = G.Winv[r * MAX_NX + c];   // the smoother whitens with inv(chol(R1)); it draws no noise
    }
  for (int r = 0; r < NX; ++r)
    for (int c = 0; c < MAX_NU; ++c) M.B[r * MAX_NU + c] = G.B[r * MAX_NU + c];
  for (int k = 0; k < 8; ++k) M.qt[k] = G.qt[k];
  M.t_switch = G.t_switch;
  M.integ_h = G.integ_h;
  M.supersample = G.supersample;
  M.nu = G.nu;
  k_smooth<NX, DYN><<<grid, SM_BLOCK, 0, stream>>>(S, M);
  return cudaGetLastError();
}

cudaError_t smooth_launch(int nx, int dyn, const SmoothP& S, const SmoothModelG& G, int grid, cudaStream_t stream) {
  if (dyn == 0) {
    switch (nx) {
      case 1: return launch_one<1, 0>(S, G, grid, stream);
      case 2: return launch_one<2, 0>(S, G, grid, stream);
      case 3: return launch_one<3, 0>(S, G, grid, stream);
      case 4: return launch_one<4, 0>(S, G, grid, stream);
      case 6: return launch_one<6, 0>(S, G, grid, stream);
      case 8: return launch_one<8, 0>(S, G, grid, stream);
      default: break;
    }
  } else if (dyn == 1 && nx == 4) {
    return launch_one<4, 1>(S, G, grid, stream);
  }
  return cudaErrorInvalidDeviceFunction;
}

}  // namespace llpf
