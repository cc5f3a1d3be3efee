// llpf_engine_list.h — the k_engine<NX, NY, DYN, RESID> instantiations of libllpf_b200.so, spread over
// LLPF_INST_GROUPS translation units (llpf_engine_inst.cu compiled once per group, in parallel: a single TU with all
// of them takes ~6 min of nvcc).  X(group, NX, NY, DYN, RESID)
//   DYN   0 linear dynamics, 1 quadtank RK4 (include/llpf.h LLPF_DYN_*)
//   RESID 0 systematic / stratified resampling, 1 residual resampling (llpf_residual.cuh), 2 Metropolis resampling
//         (llpf_metropolis.cuh; extension, a few instantiations — user-defined models get theirs at run time)
#pragma once
#define LLPF_INST_GROUPS 8
#ifdef LLPF_DISPATCH_MIN   /* quick tuning builds: only the headline instantiation */
#define LLPF_ENGINE_LIST(X) X(0, 4, 2, 0, 0)
#else
#define LLPF_ENGINE_LIST(X)                                                              \
  X(0, 4, 2, 0, 0) X(0, 1, 1, 0, 0) X(0, 2, 1, 0, 1) X(0, 8, 4, 0, 1)                    \
  X(1, 4, 2, 1, 0) X(1, 2, 1, 0, 0) X(1, 2, 2, 0, 1) X(1, 8, 2, 0, 1)                    \
  X(2, 8, 4, 0, 0) X(2, 2, 2, 0, 0) X(2, 3, 1, 0, 1) X(2, 1, 1, 0, 1)                    \
  X(3, 8, 2, 0, 0) X(3, 3, 1, 0, 0) X(3, 3, 2, 0, 1) X(3, 4, 2, 1, 1)                    \
  X(4, 6, 3, 0, 0) X(4, 3, 2, 0, 0) X(4, 3, 3, 0, 1) X(4, 4, 2, 0, 1)                    \
  X(5, 6, 2, 0, 0) X(5, 3, 3, 0, 0) X(5, 4, 1, 0, 1) X(5, 6, 3, 0, 1)                    \
  X(6, 4, 4, 0, 0) X(6, 4, 1, 0, 0) X(6, 4, 3, 0, 1) X(6, 6, 2, 0, 1)                    \
  X(7, 4, 3, 0, 0) X(7, 4, 4, 0, 1)                                                      \
  X(1, 4, 2, 0, 2) X(3, 2, 1, 0, 2) X(5, 2, 2, 0, 2) X(7, 4, 2, 1, 2)
#endif
