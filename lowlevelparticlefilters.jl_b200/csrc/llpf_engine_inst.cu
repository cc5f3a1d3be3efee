// llpf_engine_inst.cu — one group of k_engine instantiations (see llpf_engine_list.h); compiled with
// -DLLPF_INST_GROUP=<g> for g in [0, LLPF_INST_GROUPS).  Exposes the kernels to llpf_api.cu as host-side function
// pointers: the launch (cudaLaunchCooperativeKernel) and the occupancy query take them as `const void*`.
#include "llpf_engine.cuh"
#include "llpf_engine_list.h"

#ifndef LLPF_INST_GROUP
#error "compile with -DLLPF_INST_GROUP=<group>"
#endif
#define LLPF_CAT2(a, b) a##b
#define LLPF_CAT(a, b) LLPF_CAT2(a, b)

namespace llpf {

const void* LLPF_CAT(engine_kernel_group, LLPF_INST_GROUP)(int nx, int ny, int dyn, int resid) {
#define X(G, NX, NY, DYN, R) LLPF_CAT(LLPF_INST_SEL_, G)(NX, NY, DYN, R)
#define LLPF_INST_PICK(NX, NY, DYN, R) \
  if (nx == NX && ny == NY && dyn == DYN && resid == R) return (const void*)k_engine<NX, NY, DYN, R>;
#define LLPF_INST_SKIP(NX, NY, DYN, R)
#if LLPF_INST_GROUP == 0
#define LLPF_INST_SEL_0 LLPF_INST_PICK
#else
#define LLPF_INST_SEL_0 LLPF_INST_SKIP
#endif
#if LLPF_INST_GROUP == 1
#define LLPF_INST_SEL_1 LLPF_INST_PICK
#else
#define LLPF_INST_SEL_1 LLPF_INST_SKIP
#endif
#if LLPF_INST_GROUP == 2
#define LLPF_INST_SEL_2 LLPF_INST_PICK
#else
#define LLPF_INST_SEL_2 LLPF_INST_SKIP
#endif
#if LLPF_INST_GROUP == 3
#define LLPF_INST_SEL_3 LLPF_INST_PICK
#else
#define LLPF_INST_SEL_3 LLPF_INST_SKIP
#endif
#if LLPF_INST_GROUP == 4
#define LLPF_INST_SEL_4 LLPF_INST_PICK
#else
#define LLPF_INST_SEL_4 LLPF_INST_SKIP
#endif
#if LLPF_INST_GROUP == 5
#define LLPF_INST_SEL_5 LLPF_INST_PICK
#else
#define LLPF_INST_SEL_5 LLPF_INST_SKIP
#endif
#if LLPF_INST_GROUP == 6
#define LLPF_INST_SEL_6 LLPF_INST_PICK
#else
#define LLPF_INST_SEL_6 LLPF_INST_SKIP
#endif
#if LLPF_INST_GROUP == 7
#define LLPF_INST_SEL_7 LLPF_INST_PICK
#else
#define LLPF_INST_SEL_7 LLPF_INST_SKIP
#endif
  LLPF_ENGINE_LIST(X)
#undef X
  return nullptr;
}

}  // namespace llpf
