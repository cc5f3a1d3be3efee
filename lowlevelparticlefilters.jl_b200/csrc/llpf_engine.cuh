// llpf_engine.cuh — the persistent, cooperative particle-filter engine for sm_100a (v2).
//
// One launch runs a whole trajectory (or one step verb): the sequential time loop of
// forward_trajectory / loglik (reference src/filtering.jl:343-384, src/smoothing.jl:227-236) never
// returns to the host.  The grid is co-resident (cooperative launch); blocks own contiguous particle
// chunks and meet at a hand-rolled grid barrier (one red.release + spin per block).
// The host compiles the call into a tiny run-length op list (EngineP::ops); the kernel has exactly one
// inlined call site per pass kind, so the filter scalars and all per-particle state live in registers.
//
// Pass structure for ParticleFilter / AdvancedParticleFilter (filtering.jl:140-168):
//     W(1)  [P(1)+W(2)] [P(2)+W(3)] ... [P(T-1)+W(T)]  P(T)
//   where W(k) = correct!(k)  (measurement_equation! PFtypes.jl:107-120 / :226-239 + logsumexp! utils.jl:18-27)
//         P(k) = predict!(k)  (shouldresample/resample resample.jl:5-36 + propagate_particles! PFtypes.jl:122-139,
//                              ext/...DistributionsExt.jl:83-93 + reset_weights! utils.jl:73-78)
//   predict!(k) and correct!(k+1) are fused into one sweep over the particles: 2*nx*8+16 B of traffic
//   per particle-step instead of the 3*nx*8+16 B of the un-fused formulation; the copyto!(xprev,x)
//   of filtering.jl:151 disappears (in-place update, ping-pong only on resample steps).
//
// AuxiliaryParticleFilter (filtering.jl:195-217): aux_step = sweep A -> scan -> sweep B.
//
// Weight normalisation is lazy: w[] keeps the un-normalised log-weights and the pair (max, log sum)
// found by the grid reduction is applied when w is next read (same two subtractions as utils.jl:20,25).
// `we` is never materialised in the loop; ESS = S^2/Q comes out of the same reduction.
//
// Resampling (resample.jl:17-61): bins = device-wide inclusive scan of we.  FAST mode scans in 2^-62
// fixed point (u64): exact, associative, monotone, independent of block/GPU partitioning.  SERIAL mode
// is one thread doing the reference's left-to-right f64 adds (bit-exact verification mode).  Indices
// come from the *source side*: particle b owns the output slots {i : bins[b-1] <= s_i < bins[b]}, whose
// first element is found in O(1) by inverting the threshold sequence (with an exact fix-up that
// evaluates s_i in the reference's arithmetic) — the same partition as the reference's two-pointer
// walk (resample.jl:26-34), with no dependent memory chain.
#pragma once
#include "llpf_rtc_compat.h"

#include "llpf_rng.cuh"

// ------------------------------------------------------------------------------------------------
// user-defined models (DYN == LLPF_DYN_USER = 2): the reference takes arbitrary closures `dynamics(x,u,p,t)` and
// `measurement_likelihood(x,u,y,p,t)` (src/PFtypes.jl:128,232).  Closures cannot cross a C-ABI, device source can: a
// translation unit that defines these two templates (or specialisations) and instantiates k_engine<NX,NY,2,RESID> gets a
// filter with those functions inlined into the sweep — compiled at run time with NVRTC by llpf_create_user (llpf_api.cu:
// compile_user_kernel; scripts/nvrtc_probe.py compiles the same source offline).  u, y: the raw vectors of the step; p: the
// parameter vector `p` of the filter; the dynamics noise L1*z is drawn by the engine (added, or handed to the add_noise hook).
// Everything below that serves DYN == 2 is compiled only with -DLLPF_USER_MODEL, so that the kernels of the shipped library
// are byte-for-byte what was measured (the extra kernel-parameter / shared-memory fields alone perturb register allocation).
// ------------------------------------------------------------------------------------------------
#ifdef LLPF_USER_MODEL
namespace llpf_user {
template <int NX>
__device__ void dynamics(double (&x)[NX], const double* u, const double* p, double t);
template <int NX>
__device__ double loglik(const double (&x)[NX], const double* u, const double* y, const double* p, double t);
#ifdef LLPF_USER_STATE_HOOKS
// Particles that carry more than a state vector (the reference's RBParticle: nonlinear state + mean and covariance of a
// per-particle Kalman filter, rbpf.jl:1-5) need two more places to run code:
//  * add_noise: how the drawn noise nz = L1*z enters the propagated particle (x = f(xprev) already; xprev is the particle
//    the step started from). Default behaviour without the hooks: x += nz (PFtypes.jl:135). rbpf.jl:205-227 also feeds the
//    noise of the nonlinear state into the linear state's mean through a gain that depends on the particle's covariance.
//  * correct_state: the part of correct! that mutates the particle after its weight was computed (rbpf.jl:252-277: the
//    Kalman measurement update of the linear sub-state); called with the same (u, y, t) as loglik, skipped with it when
//    the measurement is missing.
template <int NX>
__device__ void add_noise(double (&x)[NX], const double (&xprev)[NX], const double (&nz)[NX], const double* u,
                          const double* p, double t);
template <int NX>
__device__ void correct_state(double (&x)[NX], const double* u, const double* y, const double* p, double t);
#endif
}  // namespace llpf_user
#endif

// Block index as the engine sees it.  The batched multi-chain kernel (llpf_engine_batch.cu, -DLLPF_BATCH) runs one
// independent single-block filter per thread block: there every block is "block 0 of a grid of one".
#ifdef LLPF_BATCH
#define LLPF_BLOCKIDX 0u
#else
#define LLPF_BLOCKIDX blockIdx.x
#endif

namespace llpf {

#ifndef LLPF_MIN_BLOCKS
#define LLPF_MIN_BLOCKS 2
#endif
#ifndef LLPF_BLOCK
#define LLPF_BLOCK 256
#endif
constexpr int BLOCK = LLPF_BLOCK;
constexpr int NWARP = BLOCK / 32;
constexpr int MAX_BLOCKS = 1024;
constexpr int MAX_NU = 8;
constexpr int MAX_NX = 8;
constexpr int MAX_WORLD = 8;
constexpr int PS = 12;  // doubles per block partial: m, s, q, sx[8], pad
constexpr int MAX_OPS = 8;
constexpr int HEAVY_T = 1024;    // offspring runs longer than this go through the grid-wide heavy-run list
constexpr int HEAVY_MAX = 2048;  // capacity of that list (entries of 3 ints); beyond it runs are written directly
constexpr int MAX_ROWS = 128;   // rows (of BLOCK particles) per block handled by the 2-barrier scan
constexpr double FIX_SCALE = 4611686018427387904.0;        // 2^62
constexpr double FIX_INV = 2.168404344971008868e-19;       // 2^-62

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// persistent scalar state of one filter (device memory; the PFstate refs maxw/t plus bookkeeping)
// ------------------------------------------------------------------------------------------------
struct Scalars {
  long long t_index;      // index(pf) = state.t[]  (PFtypes.jl:16, filtering.jl:13,152)
  long long resample_count;
  int cur;                // which x buffer holds particles(pf)
  int uniform;            // 0: w[] materialised; 1: w == -log(N) (reset!, filtering.jl:11); 2: w == log(1/N) (reset_weights!, utils.jl:75)
  int pend;               // w[] un-normalised; logical w = (w - pend_m) - pend_ls
  int stats_ahead;        // APF: (pend_m,pend_ls,...) describe raw w[] but correct! has not been called yet
  int stats_valid;        // ess valid for the current weights
  int j_identity;         // state.j == 1:N (filtering.jl:148)
  int nonfinite;          // a log-likelihood increment was not finite
  int last_resampled;
  double pend_m, pend_ls, pend_s;
  double ess;
  double ll_last;
  double ll_total;
  double bins_total;
  double xhat[MAX_NX];
  unsigned long long xseq;   // multi-GPU: number of peer exchanges done so far (identical on every rank)
};

// run-length op list: for c in [0,count): run kind with (a0 + c*da, b0 + c*db)
enum { OP_PF = 0, OP_AUX_STEP = 1, OP_AUX_CSTATS = 2, OP_FLUSH_WHIST = 3 };
enum { OPF_SKIP_MEAS = 1, OPF_RAW_WEIGHTS = 2, OPF_POST_CSTATS = 4 };
struct OpRun {
  int kind, a0, b0, count, da, db, flags, pad;
};

struct EngineP {
  double* x[2];           // SoA: component d of local particle i at x[buf][d*ld + i]
  long long ld;
  double* w;              // [n] log-weights (lazy-normalised)
  double* lam;            // [n] APF lambda (the reference aliases state.we, filtering.jl:200)
  double* bins;           // [n] cumulative weights of the last resample (state.bins)
  u64* loc;               // [n] scratch: block-local fixed-point prefix of the running scan
  int* j;                 // [n] 0-based global ancestor indices of the last resample (state.j)
  unsigned int* bar;      // grid barrier counter (zeroed by the host before every launch)
  double* partials;       // [MAX_BLOCKS*PS]
  u64* tots;              // [MAX_BLOCKS]
  Scalars* sc;
  const double* u;        // [T][nu]
  const double* y;        // [T][ny]
  double* ll_steps;       // optional per-step outputs (device)
  double* ess_steps;
  int* resampled;
  double* xhat;
  double* x_hist;         // [T][N][nx] AoS
  double* w_hist;         // [T][N]
  double* we_hist;        // [T][N]
  long long N;            // global particle count
  int n;                  // local particle count
  int first;              // global index of local particle 0
  int nops;
  int filter;             // LLPF_FILTER_*
  OpRun ops[MAX_OPS];
  int time_conv;          // 0: t=(k-1)*Ts (filtering.jl:352) ; 1: t=k*Ts (filtering.jl:181 with index from 1)
  int use_t_override;
  double t_override;
  double Ts;
  double thr;             // resample_threshold
  int strategy;           // LLPF_RESAMPLE_*
  int scan_mode;          // LLPF_SCAN_*
  int want_xhat;
  int nblocks;            // == gridDim.x
  int chunk;              // particles per block
  int rank, world;
  RngKey key;
  double fix_scale, fix_inv;  // fixed-point scan scale: 2^62 / 2^-62 for normalised weights
  int dbg_T;                  // number of time steps of the run (layout of dbg)
  long long* dbg;             // optional phase timestamps [pass][16] (block 0, thread 0; -DLLPF_PHASE_TIMING)
  // ---- sharded filters (world > 1): IPC-mapped views of every rank's arrays, index = rank ----
  double* peer_x[MAX_WORLD][2];   // particle buffers (gather source on resample steps)
  int* peer_j[MAX_WORLD];         // ancestor arrays (offspring indices are scattered to the slot's owner)
  double* peer_mbox[MAX_WORLD];   // mailboxes: [2 parities][MAX_WORLD senders][MBOX_WORDS] tagged words
  double* bcast;                  // local re-broadcast of a finished exchange: [2][MAX_WORLD][MBOX_DOUBLES]
  u64* bcast_flag;                // [2]
  // heavy offspring runs of the current resample: [0] = count, then (first slot, length, id) triples
  int* heavy;
  int* peer_heavy[MAX_WORLD];
  u64* tots2;             // [MAX_BLOCKS] second per-block total (residual resampling: the integer offspring counts)
  // ---- packed particle exchange of sharded resampling (world > 1) ----
  // A particle whose offspring run reaches into another rank's slots is SHIPPED there once per destination as a packed
  // entry {state, global id, first slot, count}: posted stores into the destination's pack_in (region of the sending
  // rank), indices from a local counter.  After the cross-GPU barrier (which carries the counts) the destination expands
  // the runs into its own j (as negative entry codes) and the gathering sweep reads the state from LOCAL memory.
  char* pack_in;                  // local: [world][pack_cap] entries of pack_stride bytes
  char* peer_pack[MAX_WORLD];     // the same area of every rank
  int* pack_cnt;                  // local: [MAX_WORLD] entries pushed to each destination in the current resample
  long long pack_cap;             // entries per (source, destination) region (= local particle count: one entry per particle at most)
  int pack_stride;                // bytes per entry: state padded to 16 B (f64 SoA particles: nx doubles; wide: 64 floats) + 16 B meta
  int pack_state_bytes;           // offset of the meta int4 {gid, first global slot, count, 0}
  int nx;                         // state dimension (f64 engines)
  int wide;                       // 1: Float32 AoS rows of 64 (llpf_wide.cuh), 0: f64 SoA
  int metro_steps;                // Metropolis resampling: proposals per slot (llpf_metropolis.cuh)
  int pad_tail;
#ifdef LLPF_USER_MODEL
  const double* user_p;   // DYN == 2: the filter's parameter vector `p` (device memory), handed to the user functions
#endif
};
constexpr int MBOX_DOUBLES = 16;  // broadcast-buffer stride per rank ([1..] payload)
constexpr int MBOX_WORDS = 32;    // mailbox slot: 2 tagged 8-byte words per payload double

#ifdef LLPF_PHASE_TIMING
#define LLPF_TS(P, sh, k)                                                              \
  do {                                                                                 \
    if ((P).dbg && blockIdx.x == 0 && threadIdx.x == 0) (P).dbg[(size_t)(sh).pass_id * 16 + (k)] = clock64(); \
  } while (0)
// every block: wall-clock (globaltimer, ns) of event k in {0,1,2,3} of the current pass; the per-block table
// follows block 0's [pass][16] table: [pass][MAX_BLOCKS][4]  (scripts/skew.py)
#define LLPF_TSB(P, sh, T_, k)                                                                       \
  do {                                                                                              \
    if ((P).dbg && threadIdx.x == 0) {                                                              \
      unsigned long long gt_;                                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                       \
      (P).dbg[(size_t)16 * ((T_) + 2) + ((size_t)(sh).pass_id * MAX_BLOCKS + blockIdx.x) * 4 + (k)] = (long long)gt_; \
    }                                                                                               \
  } while (0)
#else
#define LLPF_TS(P, sh, k) do { } while (0)
#define LLPF_TSB(P, sh, T_, k) do { } while (0)
#endif

template <int NX, int NY>
struct ModelP {
  double A[NX * NX];      // row-major
  double L1[NX * NX];     // row-major, lower Cholesky factor of R1 (noise = L1*z, utils.jl:260-268)
  double G[NY * NX];      // row-major: inv(chol(R2)) * C      (whitened measurement matrix)
  double W[NY * NY];      // row-major, lower: inv(chol(R2))
  double B[NX * MAX_NU];  // row-major NX x nu
  double c0;              // mvnormal_c0 = -(ny*log(2pi) + logdet R2)/2   (utils.jl:254-257)
  double qt[8];           // quadtank: {-a/A, -(a*f)/A, a/A, 2g, g1k1/A, g2k2/A, (1-g2)k2/A, (1-g1)k1/A}
  double t_switch;
  double integ_h;         // Ts0 / supersample  (utils.jl:223)
  int supersample;
  int nu;
};

// ------------------------------------------------------------------------------------------------
// synchronisation + block collectives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All blocks are co-resident (cooperative launch).  One arrive + spin per block; the spin polls with
// relaxed loads and issues a single acquire once the count is reached (one L1 invalidation per barrier).
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    red_release_add_u32(bar, 1u);
    while ((int)(ld_relaxed_u32(bar) - target) < 0) {
    }
    (void)ld_acquire_u32(bar);
  }
  __syncthreads();
}

__host__ __device__ constexpr int mdl_stride(int nx) { return (nx + 1) & ~1; }   // even row stride: rows are 16-byte aligned

// ------------------------------------------------------------------------------------------------
// cross-GPU exchange over NVLink peer memory (one process per GPU, arrays mapped with CUDA IPC)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 ld_acquire_sys_u64(const u64* p) {
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(u64* p, u64 v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// release fence at system scope: what a message-passing pattern over NVLink needs (`__threadfence_system()` is the
// sequentially-consistent fence.sc.sys, measured at ~5 us per use on the critical path of a sharded resample)
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// Packed particle exchange, release side (tuning switch; 3 is what ships):
//   1: every pushing thread fences, block 0 posts the count with a relaxed store
//   2: no per-thread fence, block 0 posts with st.release.sys (relies on cumulativity through the grid barrier alone;
//      2.5 us per resample step faster than 3 on 2 GPUs, bit-identical in the parity runs; not shipped: 3 also holds if a
//      pushing SM's posted writes were not covered by another SM's release)
//   3: both
#ifndef LLPF_XCHG_VARIANT
#define LLPF_XCHG_VARIANT 3
#endif
__device__ __forceinline__ void st_relaxed_sys_u64(u64* p, u64 v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ u64 ld_relaxed_sys_u64(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_relaxed_gpu_u64(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_acquire_gpu_u64(const u64* p) {
  u64 v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u64(u64* p, u64 v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct Shared {
  // model matrices, read with volatile 16-byte shared loads inside the particle loop (see llpf_math.cuh:
  // the compiler would otherwise hoist ~35 loop-invariant constant loads into registers and spill them)
  alignas(16) double mA[MAX_NX * MAX_NX];   // row r at mA + r*stride
  alignas(16) double mL[MAX_NX * MAX_NX];   // lower Cholesky factor of R1
  alignas(16) double mG[8 * MAX_NX];        // whitened measurement matrix
  double red_a[NWARP];
  double red_b[NWARP * (2 + MAX_NX)];
  u64 wtot[NWARP];
  u64 wt[MAX_ROWS * NWARP];     // per (row, warp) totals -> exclusive offsets of the block-local scan
  int wtf[MAX_ROWS * NWARP];    // F(bins) at the segment starts (first output slot of each segment)
  double bu[MAX_NX];
  double yt[8];
  int skip;
  int pass_id;
  int is_last;                  // reduce_stats: this block arrived last and runs the combine
  uint32_t stat_words[32];      // reduce_stats: payload halves of the finished statistics
  double peer_vals[MAX_WORLD * (3 + MAX_NX)];   // reduce_stats: the ranks' statistics (sharded filters)
  u64 offs[MAX_BLOCKS + 1];     // exclusive block offsets of the fixed-point scan
  alignas(16) MathTab mt;       // log / exp tables + polynomial coefficients of llpf_math.cuh
#ifdef LLPF_USER_MODEL
  // DYN == 2 (user-defined model): the raw vectors of the pass
  double u_prop[MAX_NU];        // u of the step being propagated
  double u_weigh[MAX_NU];       // u of the step being weighed
  double y_raw[8];
  const double* user_p;
#endif
};

// All-gather of NV doubles per rank.  Every block of every rank calls it with identical `mine` (the
// values every block derived from the same local data after a local grid barrier).
//   block 0, thread r (r != rank): writes the rank's values into peer r's mailbox — "LL" protocol as in NCCL's
//     low-latency path: every 8-byte word carries 4 bytes of payload and the 4-byte sequence number, 8-byte
//     stores are single NVLink transactions, so a word is valid exactly when its tag equals the expected
//     sequence number and no fence / flag is needed;
//   EVERY block: polls the words the peers wrote into the LOCAL mailbox (one thread per word).  (v6 had block 0
//     poll and re-broadcast through a flag: one more fence + L2 hop on the critical path of every exchange.)
// Slots are double-buffered by the parity of the sequence number; a rank can never be more than one
// exchange ahead of a peer (it needs that peer's post to get past the current one), and a post is only
// made after a local grid-wide synchronisation, i.e. after all of the rank's blocks finished reading the
// previous exchange of the same parity.
template <int NV>
__device__ __forceinline__ void peer_allgather(const EngineP& P, Shared& sh, u64& xseq, const double (&mine)[NV],
                                               double (&all)[MAX_WORLD][NV]) {
  static_assert(2 * NV <= MBOX_WORDS && NV <= 3 + MAX_NX, "payload too large");
  xseq += 1;
  const int par = (int)(xseq & 1ull);
  const u64 tag = (xseq & 0xffffffffull) << 32;
  if (LLPF_BLOCKIDX == 0 && threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
    const int r = threadIdx.x;
    u64* out = reinterpret_cast<u64*>(P.peer_mbox[r]) + ((size_t)par * MAX_WORLD + P.rank) * MBOX_WORDS;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const u64 bits = (u64)__double_as_longlong(mine[k]);
      st_relaxed_sys_u64(out + 2 * k, tag | (bits & 0xffffffffull));
      st_relaxed_sys_u64(out + 2 * k + 1, tag | (bits >> 32));
    }
  }
  uint32_t* words = reinterpret_cast<uint32_t*>(sh.peer_vals);   // [world][2*NV] payload halves
  __syncthreads();                                               // earlier readers of sh.peer_vals are done
  if ((int)threadIdx.x < P.world * 2 * NV) {
    const int r = threadIdx.x / (2 * NV), wd = threadIdx.x % (2 * NV);
    if (r != P.rank) {
      const u64* in = reinterpret_cast<const u64*>(P.peer_mbox[P.rank]) + ((size_t)par * MAX_WORLD + r) * MBOX_WORDS + wd;
      u64 v;
      do { v = ld_relaxed_sys_u64(in); } while ((v & 0xffffffff00000000ull) != tag);
      words[threadIdx.x] = (uint32_t)v;
    }
  }
  __syncthreads();
  for (int r = 0; r < P.world; ++r) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
      all[r][k] = (r == P.rank) ? mine[k]
                                : __hiloint2double((int)words[r * 2 * NV + 2 * k + 1], (int)words[r * 2 * NV + 2 * k]);
  }
}

// cross-GPU barrier carrying one value per rank (the length of the rank's heavy-run list).  Call right after a
// local grid barrier; every thread that stored into peer memory must have executed a system-scope release fence
// before arriving at that barrier, so all of this rank's peer stores are performed before block 0 posts.
__device__ __forceinline__ void peer_barrier(const EngineP& P, Shared& sh, u64& xseq, double payload,
                                             double (&all)[MAX_WORLD][1]) {
  const double mine[1] = {payload};
  peer_allgather<1>(P, sh, xseq, mine, all);
}

template <int K>
__device__ __forceinline__ void lds_row(const double* row, double (&out)[K]) {   // row is 16-byte aligned, padded
#pragma unroll
  for (int c = 0; c < K; c += 2) {
    const double2 v = lds2v(reinterpret_cast<const double2*>(row + c));
    out[c] = v.x;
    if (c + 1 < K) out[c + 1] = v.y;
  }
}
template <int NX, int NY>
struct ModelP;
template <int NX, int NY>
__device__ __forceinline__ void model_to_shared(const ModelP<NX, NY>& M, Shared& sh);

// max over the block, result in every thread (deterministic).  One barrier: red_a and red_b are
// used alternately (max -> sum -> max -> sum), which orders every reuse behind a barrier.
__device__ __forceinline__ double block_max(double v, Shared& sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh.red_a[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = sh.red_a[0];
#pragma unroll
  for (int i = 1; i < NWARP; ++i) r = fmax(r, sh.red_a[i]);
  return r;
}

// sums of M values over the block, results in every thread (fixed order => bitwise identical everywhere)
template <int M>
__device__ __forceinline__ void block_sum(double (&v)[M], Shared& sh) {
#pragma unroll
  for (int k = 0; k < M; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < M; ++k) sh.red_b[(threadIdx.x >> 5) * M + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < M; ++k) {
    double r = sh.red_b[k];
#pragma unroll
    for (int i = 1; i < NWARP; ++i) r += sh.red_b[i * M + k];
    v[k] = r;
  }
}

// online (max, sum exp, sum exp^2 [, sum exp*x]) accumulator: exactly one exp per sample
template <int NX>
struct Online {
  double m, s, q;
  double sx[NX];
  __device__ __forceinline__ void init() {
    m = -DBL_MAX; s = 0.0; q = 0.0;
#pragma unroll
    for (int d = 0; d < NX; ++d) sx[d] = 0.0;
  }
  __device__ __forceinline__ void add(double wv, const double (&x)[NX], bool with_x, const MathTab& T) {
    const double d = wv - m;
    const double e = exp_nonpos(-fabs(d), T);
    if (d > 0.0) {   // new running maximum: rescale what we have
      s = fma(s, e, 1.0);
      q = fma(q, e * e, 1.0);
      if (with_x) {
#pragma unroll
        for (int k = 0; k < NX; ++k) sx[k] = fma(sx[k], e, x[k]);
      }
      m = wv;
    } else {
      s += e;
      q = fma(e, e, q);
      if (with_x) {
#pragma unroll
        for (int k = 0; k < NX; ++k) sx[k] = fma(e, x[k], sx[k]);
      }
    }
  }
};

struct Stats {
  double m, s, q;
  double sx[MAX_NX];
};

// Layout of the synchronisation words behind EngineP::bar (zeroed by the host before every launch):
//   bar[0]            grid-barrier arrival counter
//   bar[32]           reduction arrival counter (its own 128-byte line)
//   bar[64 .. 191]    64 tagged 8-byte words: the finished statistics of a reduction, [2 parities][32 words]
constexpr int BAR_RED_CNT = 32;
constexpr int BAR_RES_WORDS = 64;      // offset in unsigned units
constexpr int BAR_TOTAL_WORDS = 192;   // unsigned units the host must allocate and zero

__device__ __forceinline__ unsigned atom_add_acq_rel_u32(unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Reduce the per-thread online accumulators to one block partial and publish it; the LAST block to
// arrive (atomic counter) combines all block partials in a fixed order — on sharded filters it also
// performs the per-step exchange with the peer GPUs — and publishes the finished statistics as tagged
// 8-byte words (4 B payload + 4 B sequence number, NCCL-LL style: a word is valid exactly when its tag
// matches, so one L2 round trip delivers the data and no flag/fence pair is needed).  Every other block
// polls those words.  All blocks end up with bitwise-identical Stats, run-to-run deterministic (the
// combine order is fixed; which block happens to run it does not matter).
// Replaces (v6) "grid barrier + every block re-reads all partials": one atomic + one polled line per block.
template <int NX>
__device__ __forceinline__ Stats reduce_stats(const EngineP& P, Shared& sh, Online<NX>& acc,
                                              bool with_x, unsigned& red_seq, u64& xseq) {
  constexpr int NV = 3 + NX;          // m, s, q, sx[NX]
  constexpr int NW = 2 * NV;          // tagged words
  static_assert(NW <= 32, "statistics do not fit the result line");
  const double mb = block_max(acc.m, sh);
  const double sc = exp_nonpos(acc.m - mb, sh.mt);  // 0 for empty threads
  double v[2 + NX];
  v[0] = acc.s * sc;
  v[1] = acc.q * (sc * sc);
#pragma unroll
  for (int d = 0; d < NX; ++d) v[2 + d] = with_x ? acc.sx[d] * sc : 0.0;
  block_sum<2 + NX>(v, sh);
  LLPF_TS(P, sh, 6);
  red_seq += 1;
  const unsigned seq = red_seq;
  if (P.world > 1) xseq += 1;
  u64* res = reinterpret_cast<u64*>(P.bar + BAR_RES_WORDS) + (size_t)(seq & 1u) * 32;
  if (threadIdx.x == 0) {   // partial = {m, s, q, sx[NX]} as 16-byte pairs
    double2* p = reinterpret_cast<double2*>(P.partials + (size_t)LLPF_BLOCKIDX * PS);
    double pk[2 * ((3 + NX + 1) / 2)];
    pk[0] = mb; pk[1] = v[0]; pk[2] = v[1];
#pragma unroll
    for (int d = 0; d < NX; ++d) pk[3 + d] = v[2 + d];
    if ((3 + NX) & 1) pk[3 + NX] = 0.0;
#pragma unroll
    for (int k = 0; k < (3 + NX + 1) / 2; ++k) __stcg(p + k, make_double2(pk[2 * k], pk[2 * k + 1]));
    // release: my partial is visible before the count; acquire: if I am last, everyone's partial is visible to me
    const unsigned prev = atom_add_acq_rel_u32(P.bar + BAR_RED_CNT, 1u);
    sh.is_last = (prev + 1u == seq * (unsigned)P.nblocks) ? 1 : 0;
  }
  __syncthreads();
  LLPF_TS(P, sh, 7);
  if (sh.is_last) {
    // combine: thread 0's acquire invalidated L1, so these (L1-allocating) loads are fresh and the second
    // sweep over the same lines hits L1.  No block overwrites its partial before it has seen the result.
    double m = -DBL_MAX;
    for (int b = threadIdx.x; b < P.nblocks; b += BLOCK)
      m = fmax(m, __ldca(reinterpret_cast<const double2*>(P.partials + (size_t)b * PS)).x);
    m = block_max(m, sh);
    double t[2 + NX];
#pragma unroll
    for (int k = 0; k < 2 + NX; ++k) t[k] = 0.0;
    for (int b = threadIdx.x; b < P.nblocks; b += BLOCK) {
      const double2* p = reinterpret_cast<const double2*>(P.partials + (size_t)b * PS);
      double pk[2 * ((3 + NX + 1) / 2)];
#pragma unroll
      for (int k = 0; k < (3 + NX + 1) / 2; ++k) {
        const double2 q2 = (with_x || k < 2) ? __ldca(p + k) : make_double2(0.0, 0.0);
        pk[2 * k] = q2.x; pk[2 * k + 1] = q2.y;
      }
      const double e = exp_nonpos(pk[0] - m, sh.mt);
      t[0] = fma(pk[1], e, t[0]);
      t[1] = fma(pk[2], e * e, t[1]);
      if (with_x) {
#pragma unroll
        for (int d = 0; d < NX; ++d) t[2 + d] = fma(pk[3 + d], e, t[2 + d]);
      }
    }
    block_sum<2 + NX>(t, sh);
    LLPF_TS(P, sh, 9);
    if (P.world > 1) {
      // one all-gather of (m, s, q, sx) per rank over NVLink (tagged words written straight into every peer's
      // mailbox, thread r talks to rank r); combined in rank order on every rank
      const int par = (int)(xseq & 1ull);
      const u64 tag = (xseq & 0xffffffffull) << 32;
      double mine[NV];
      mine[0] = m;
#pragma unroll
      for (int k = 0; k < 2 + NX; ++k) mine[1 + k] = t[k];
      // one thread per (peer, word): every store and every poll is a single independent transaction (the first version
      // had thread r write and then poll rank r's 2*NV words one after the other: ~2*NV dependent L2 round trips)
      static_assert(MAX_WORLD * NW <= BLOCK, "one thread per (peer, word)");
      uint32_t* pw = reinterpret_cast<uint32_t*>(sh.peer_vals);   // [world][NW] payload halves
      if ((int)threadIdx.x < P.world * NW) {
        const int r = threadIdx.x / NW, wd = threadIdx.x % NW;
        double val = mine[0];
#pragma unroll
        for (int k = 1; k < NV; ++k)
          if ((wd >> 1) == k) val = mine[k];
        const u64 bits = (u64)__double_as_longlong(val);
        const uint32_t half = (wd & 1) ? (uint32_t)(bits >> 32) : (uint32_t)bits;
        if (r != P.rank) {
          u64* out = reinterpret_cast<u64*>(P.peer_mbox[r]) + ((size_t)par * MAX_WORLD + P.rank) * MBOX_WORDS;
          st_relaxed_sys_u64(out + wd, tag | (u64)half);
          const u64* in = reinterpret_cast<const u64*>(P.peer_mbox[P.rank]) + ((size_t)par * MAX_WORLD + r) * MBOX_WORDS;
          u64 v;
          do { v = ld_relaxed_sys_u64(in + wd); } while ((v & 0xffffffff00000000ull) != tag);
          pw[r * NW + wd] = (uint32_t)v;
        } else {
          pw[r * NW + wd] = half;
        }
      }
      __syncthreads();
      double gm = -DBL_MAX;
      for (int r = 0; r < P.world; ++r) gm = fmax(gm, __hiloint2double((int)pw[r * NW + 1], (int)pw[r * NW]));
#pragma unroll
      for (int k = 0; k < 2 + NX; ++k) t[k] = 0.0;
      for (int r = 0; r < P.world; ++r) {
        double pv[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) pv[k] = __hiloint2double((int)pw[r * NW + 2 * k + 1], (int)pw[r * NW + 2 * k]);
        const double e = exp_nonpos(pv[0] - gm, sh.mt);
        t[0] = fma(pv[1], e, t[0]);
        t[1] = fma(pv[2], e * e, t[1]);
#pragma unroll
        for (int d = 0; d < NX; ++d) t[2 + d] = fma(pv[3 + d], e, t[2 + d]);
      }
      m = gm;
    }
    if (threadIdx.x < NW) {   // publish: word 2k = low half of value k, word 2k+1 = high half
      double val = m;
#pragma unroll
      for (int k = 0; k < 2 + NX; ++k)
        if ((int)(threadIdx.x >> 1) == 1 + k) val = t[k];
      const u64 bits = (u64)__double_as_longlong(val);
      const uint32_t half = (threadIdx.x & 1) ? (uint32_t)(bits >> 32) : (uint32_t)bits;
      sh.stat_words[threadIdx.x] = half;
      st_relaxed_gpu_u64(res + threadIdx.x, ((u64)seq << 32) | half);
    }
  } else if (threadIdx.x < NW) {
    u64 wv;
    do { wv = ld_relaxed_gpu_u64(res + threadIdx.x); } while ((uint32_t)(wv >> 32) != seq);
    sh.stat_words[threadIdx.x] = (uint32_t)wv;
  }
  __syncthreads();
  LLPF_TS(P, sh, 8);
  Stats st;
  double out[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k)
    out[k] = __hiloint2double((int)sh.stat_words[2 * k + 1], (int)sh.stat_words[2 * k]);
  st.m = out[0]; st.s = out[1]; st.q = out[2];
#pragma unroll
  for (int d = 0; d < NX; ++d) st.sx[d] = out[3 + d];
  return st;
}

// ------------------------------------------------------------------------------------------------
// scan + source-side index construction (shared by the engine and the stand-alone resample entry)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 to_fixed(double we, double scale) {
  // we*scale in [0,2^62] (normalised weights: scale = 2^62); NaN/negative -> 0
  const double v = we * scale;
  return (v > 0.0) ? __double2ull_rn(fmin(v, FIX_SCALE)) : 0ull;
}

// Stage 1: block-local scan of the block's chunk [beg,end).
// FAST  : loc[i] = inclusive fixed-point prefix within its (row, warp) segment; sh.wt[row*NWARP+warp] = exclusive
//         offset of that segment inside the block (two block barriers in total); tots[b] = block total.
//         (Chunks of more than MAX_ROWS rows fall back to a per-row barrier and keep the full prefix in loc[].)
// SERIAL: bins[i] = we_i (double); the single-thread pass runs after the grid barrier.
// Returns true when the (row, warp) table is in use.
template <class LoadFn, class WeFn>
__device__ __forceinline__ bool scan_stage1(const EngineP& P, Shared& sh, int beg, int end, LoadFn loadfn, WeFn wefn) {
  if (P.scan_mode != 0) {
    for (int i = beg + threadIdx.x; i < end; i += BLOCK) __stcg(P.bins + i, wefn(i, loadfn(i)));
    return false;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = (end - beg + BLOCK - 1) / BLOCK;
  if (rows <= MAX_ROWS) {
    // software pipeline: the raw weight of the next row is in flight while this row is scanned
    double raw_next = (beg + (int)threadIdx.x < end) ? loadfn(beg + threadIdx.x) : 0.0;
    for (int k = 0; k < rows; ++k) {
      const int i = beg + k * BLOCK + threadIdx.x;
      const double raw = raw_next;
      if (i + BLOCK < end) raw_next = loadfn(i + BLOCK);
      u64 v = (i < end) ? to_fixed(wefn(i, raw), P.fix_scale) : 0ull;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (i < end) __stcg(P.loc + i, v);
      if (lane == 31) sh.wt[k * NWARP + warp] = v;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the rows*NWARP segment totals, in (row, warp) order
      const int n = rows * NWARP;
      const int per = (n + 31) >> 5;
      const int s0 = lane * per, s1 = (s0 + per < n) ? s0 + per : n;
      u64 sum = 0;
      for (int q = s0; q < s1; ++q) sum += sh.wt[q];
      u64 incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      u64 run = incl - sum;
      for (int q = s0; q < s1; ++q) {
        const u64 t = sh.wt[q];
        sh.wt[q] = run;
        run += t;
      }
      if (lane == 31) __stcg(P.tots + LLPF_BLOCKIDX, incl);
    }
    __syncthreads();
    return true;
  }
  u64 carry = 0;
  for (int base = beg; base < end; base += BLOCK) {
    const int i = base + threadIdx.x;
    u64 v = (i < end) ? to_fixed(wefn(i, loadfn(i)), P.fix_scale) : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) sh.wtot[warp] = v;
    __syncthreads();
    u64 woff = 0, rtot = 0;
#pragma unroll
    for (int k = 0; k < NWARP; ++k) {
      const u64 t = sh.wtot[k];
      if (k < warp) woff += t;
      rtot += t;
    }
    if (i < end) __stcg(P.loc + i, carry + woff + v);
    carry += rtot;
  }
  if (threadIdx.x == 0) __stcg(P.tots + LLPF_BLOCKIDX, carry);
  return false;
}

// After the grid barrier that follows stage 1 (FAST): exclusive block offsets into sh.offs[0..nb]
// (exact integer arithmetic: any summation order gives the same bits).
__device__ __forceinline__ void scan_block_offsets_from(const u64* tots, int nb, Shared& sh, u64 base_fixed) {
  constexpr int PER = (MAX_BLOCKS + BLOCK - 1) / BLOCK;
  u64 t4[PER];
  u64 mine = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int b = threadIdx.x * PER + k;
    t4[k] = (b < nb) ? __ldcg(tots + b) : 0ull;
    mine += t4[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 v = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u64 t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();
  if (lane == 31) sh.wtot[warp] = v;
  __syncthreads();
  u64 woff = 0;
#pragma unroll
  for (int k = 0; k < NWARP; ++k)
    if (k < warp) woff += sh.wtot[k];
  u64 excl = base_fixed + woff + v - mine;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int b = threadIdx.x * PER + k;
    if (b <= nb) sh.offs[b] = excl;
    excl += t4[k];
  }
  __syncthreads();
}
__device__ __forceinline__ void scan_block_offsets(const EngineP& P, Shared& sh, u64 base_fixed) {
  scan_block_offsets_from(P.tots, P.nblocks, sh, base_fixed);
}

// SERIAL mode: the reference's cumsum (resample.jl:19-22), one thread, strict left-to-right f64 adds.
static __device__ __noinline__ void scan_serial(double* bins, int n) {
  double acc = __ldcg(bins);
  int i = 1;
  for (; i + 8 <= n; i += 8) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(bins + i + k);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc = __dadd_rn(acc, v[k]);
      __stcg(bins + i + k, acc);
    }
  }
  for (; i < n; ++i) {
    acc = __dadd_rn(acc, __ldcg(bins + i));
    __stcg(bins + i, acc);
  }
}

struct Thresholds {
  double r, step, total, M;
  int Mi;
  int strategy;
  uint32_t step_idx;
  const double* u_slots;   // stratified: caller-supplied rand() per slot (stand-alone entry), else RNG
  // Julia's "nice rational" range (llpf_julia_range.h; stand-alone entry only): r = ref.hi, step = step.hi and
  int twice;               // 1: elements are the double-double TwicePrecision getindex below
  int offset;              // 1-based index of the reference element
  double ref_lo, step_lo;
  // integer thresholds (FAST scan, systematic, power-of-two M, 2^62 scale): s_i = r + i/M exactly, in 2^-62 fixed point
  int ipath;               // 1: F(V) = ceil((V - Rf) / 2^ish) on the fixed-point prefix V itself
  int ish;                 // 62 - log2(M)
  u64 Rf;                  // floor(r * 2^62)
};

// the Float64 range of resample.jl:24 as the host resolved it (llpf_julia_range.h): nullptr / rational == 0 -> literal path
struct RangeArg {
  double ref_hi, ref_lo, step_hi, step_lo;
  int offset, rational;
};

// element gi (0-based) of the systematic threshold range.  Literal path: s = fl(r + fl(gi * fl(1/M))) — the product and
// the sum rounded separately (no FMA).  Rational path: Julia's unsafe_getindex for TwicePrecision ranges,
//   u = i - offset ; (x_hi, x_lo) = add12(ref.hi, u*step.hi) ; s = x_hi + (x_lo + (u*step.lo + ref.lo))
__device__ __forceinline__ double systematic_threshold(const Thresholds& th, int gi) {
  if (th.twice) {
    const double u = (double)(gi + 1 - th.offset);
    const double shift_hi = __dmul_rn(u, th.step), shift_lo = __dmul_rn(u, th.step_lo);
    double big = th.r, little = shift_hi;
    if (fabs(little) > fabs(big)) { big = shift_hi; little = th.r; }
    const double x_hi = __dadd_rn(big, little);
    const double x_lo = __dadd_rn(__dsub_rn(big, x_hi), little);
    return __dadd_rn(x_hi, __dadd_rn(x_lo, __dadd_rn(shift_lo, th.ref_lo)));
  }
  return __dadd_rn(th.r, __dmul_rn((double)gi, th.step));
}

// s[i]: systematic_threshold above (Julia StepRangeLen getindex; resample.jl:24).
// Stratified: ((i + rand_i)/M)*bins[N] (resample.jl:49).
__device__ __forceinline__ double threshold(const Thresholds& th, const RngKey& key, int gi) {
  if (th.strategy == 1) {
    double u;
    if (th.u_slots) {
      u = __ldcg(th.u_slots + gi);
    } else {
      const uint4 r = rng_block(key, ST_STRAT, th.step_idx, (unsigned long long)(unsigned)gi, 0);
      u = uniform53(r.x, r.y);
    }
    return __dmul_rn(__ddiv_rn(__dadd_rn((double)gi, u), th.M), th.total);
  }
  return systematic_threshold(th, gi);
}

__device__ __forceinline__ Thresholds make_thresholds(const EngineP& P, double total, double u01, int Mslots,
                                                      uint32_t step_idx, const double* u_slots) {
  Thresholds th;
  th.total = total;
  th.M = (double)Mslots;
  th.Mi = Mslots;
  th.strategy = P.strategy;
  th.step_idx = step_idx;
  th.u_slots = u_slots;
  th.step = __ddiv_rn(1.0, th.M);
  // r = rand()*bins[end]/N   (resample.jl:23; note /N, N = length(we))
  th.r = __ddiv_rn(__dmul_rn(u01, total), (double)P.N);
  th.twice = 0; th.offset = 1; th.ref_lo = 0.0; th.step_lo = 0.0;
  // The fixed-point scan knows every cumulative weight as an exact integer V (units of 2^-62).  For M = 2^k slots the
  // thresholds r + i/M are exact in the same units up to the fraction of r*2^62, which cannot change an integer
  // comparison: s_i >= V  <=>  i >= ceil((V - floor(r 2^62)) / 2^(62-k)).  A shift instead of ~25 FP64 / conversion
  // instructions per evaluation.  It compares against the EXACT r + i/M where the reference rounds the sum to 53 bits:
  // the two can differ only for a cumulative weight within 2^-53 of a threshold (~1e-4 per resample at N = 2^20, two
  // orders of magnitude below the FAST scan's own rounding distance to the serial cumsum).
  th.ipath = 0; th.ish = 0; th.Rf = 0ull;
  if (P.strategy == 0 && P.scan_mode == 0 && u_slots == nullptr && P.fix_scale == FIX_SCALE && Mslots > 0 &&
      (Mslots & (Mslots - 1)) == 0) {
    th.ipath = 1;
    th.ish = 62 - (31 - __clz(Mslots));
    th.Rf = __double2ull_rz(th.r * FIX_SCALE);
  }
  return th;
}
__device__ __forceinline__ int first_slot_ge_fixed(const Thresholds& th, u64 V) {
  if (V <= th.Rf) return 0;
  const u64 q = (V - th.Rf + ((1ull << th.ish) - 1ull)) >> th.ish;
  return (int)(q < (u64)th.Mi ? q : (u64)th.Mi);
}
__device__ __forceinline__ double resample_u01(const RngKey& key, uint32_t step_idx) {
  const uint4 r = rng_block(key, ST_RESAMPLE, step_idx, 0ull, 0);
  return uniform53(r.x, r.y);
}

// F(v) = min{ i in [0,M] : s_i >= v }  (M if none).  The thresholds are nondecreasing in i, so an
// estimate from the real-valued inverse plus an exact fix-up (evaluating s_i exactly as the reference
// does) gives the same partition of the slots as the reference's comparison `s[i] < bins[b]`.
__device__ __forceinline__ int first_slot_ge(const Thresholds& th, const RngKey& key, double v) {
  if (th.strategy != 1) {
    // systematic: s_i = fl(r + fl(i*step)); estimate ceil((v - r) M), then check the two neighbours once
    int i0 = __double2int_ru((v - th.r) * th.M);
    i0 = max(0, min(i0, th.Mi));
    const double s_m = systematic_threshold(th, i0 - 1);   // s[i0-1]
    const double s_0 = systematic_threshold(th, i0);       // s[i0]
    const bool down = (i0 > 0) && (s_m >= v);
    const bool up = (i0 < th.Mi) && (s_0 < v);
    if (!(down || up)) return i0;          // the estimate is exact (always, up to rounding ties)
    if (down) {
      --i0;
      while (i0 > 0 && threshold(th, key, i0 - 1) >= v) --i0;
    } else {
      ++i0;
      while (i0 < th.Mi && threshold(th, key, i0) < v) ++i0;
    }
    return i0;
  }
  const double t = (v / th.total) * th.M - 1.0;
  int i0;
  if (!(t > 0.0)) i0 = 0;
  else if (t >= th.M) i0 = th.Mi;
  else i0 = (int)ceil(t);
  while (i0 > 0 && threshold(th, key, i0 - 1) >= v) --i0;
  while (i0 < th.Mi && threshold(th, key, i0) < v) ++i0;
  return i0;
}

// Destination of a global output slot: j[0] is global slot `base` (sharded filters: the rank's first slot; runs are
// clipped to the rank's own slots before they get here, the rest travels as packed entries — push_remote_parts).
template <class T>
struct SlotRouter {
  T* j;
  int base;
  int* heavy;           // grid-wide list for very long runs (nullptr: write everything directly)
  __device__ __forceinline__ T* at(int slot) const { return j + (slot - base); }
};

// write `id` into slots [lo, lo+cnt): runs of up to 4 by predicated stores of the owning lane, longer
// runs by the whole warp.  Must be called by all 32 lanes (cnt = 0 for idle lanes).
template <class T>
__device__ __forceinline__ void scatter_runs(const SlotRouter<T>& R, int lo, int cnt, T id) {
  const int lane = threadIdx.x & 31;
  // A particle that owns thousands of output slots (degenerate weights: the normal state of a high-dimensional
  // filter) would serialise the whole resample on one warp.  Such runs are queued instead; after the barrier
  // that closes the scatter every block fills the part of each queued run that falls into its own slots.
  if (R.heavy != nullptr && cnt > HEAVY_T) {
    const int idx = atomicAdd(R.heavy, 1);
    if (idx < HEAVY_MAX) {
      __stcg(R.heavy + 1 + 3 * idx, lo);
      __stcg(R.heavy + 2 + 3 * idx, cnt);
      __stcg(R.heavy + 3 + 3 * idx, (int)id);
      cnt = 0;
    }
  }
  if (cnt <= 4) {
    if (cnt > 0) __stcg(R.at(lo), id);
    if (cnt > 1) __stcg(R.at(lo + 1), id);
    if (cnt > 2) __stcg(R.at(lo + 2), id);
    if (cnt > 3) __stcg(R.at(lo + 3), id);
  }
  unsigned heavy = __ballot_sync(0xffffffffu, cnt > 4);
  while (heavy) {
    const int src = __ffs(heavy) - 1;
    heavy &= heavy - 1;
    const int lo_s = __shfl_sync(0xffffffffu, lo, src);
    const int c = __shfl_sync(0xffffffffu, cnt, src);
    const T id_s = __shfl_sync(0xffffffffu, id, src);
    for (int s = lane; s < c; s += 32) __stcg(R.at(lo_s + s), id_s);
  }
}

// ---- packed particle exchange (sharded filters) ---------------------------------------------------------------------
// Source side.  Must be called by all 32 lanes.  The part of local particle `li`'s offspring run [lo, lo+cnt) (GLOBAL
// slots) that lies in another rank's slot range is shipped to that rank as ONE packed entry {state, gid, first slot,
// count} — posted stores into the destination's pack_in, region of this rank, index from a local counter (one atomic per
// group of lanes with the same destination).  Returns the run clipped to this rank's own slots in (lo, cnt).
__device__ __forceinline__ void push_remote_parts(const EngineP& P, int pack_buf, int& lo, int& cnt, int gid, int li,
                                                  int& pushed) {
  const int own_lo = P.first, own_hi = P.first + P.n;
  const int hi = lo + cnt;
  const bool rem = (cnt > 0) && (lo < own_lo || hi > own_hi);
  if (!__any_sync(0xffffffffu, rem)) return;
  const int lane = threadIdx.x & 31;
  int d = rem ? lo / P.n : 0;
  const int d_last = rem ? (hi - 1) / P.n : -1;
  for (;;) {
    if (d == P.rank) ++d;   // the own part stays with the caller
    const bool act = rem && d <= d_last;
    if (!__any_sync(0xffffffffu, act)) break;
    const unsigned grp = __match_any_sync(0xffffffffu, act ? d : (MAX_WORLD + lane));
    if (act) {
      const int leader = __ffs(grp) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(P.pack_cnt + d, __popc(grp));
      base = __shfl_sync(grp, base, leader);
      const int idx = base + __popc(grp & ((1u << lane) - 1u));
      const int first = max(lo, d * P.n);
      const int c = min(hi, (d + 1) * P.n) - first;
      char* e = P.peer_pack[d] + ((size_t)P.rank * (size_t)P.pack_cap + (size_t)idx) * (size_t)P.pack_stride;
      if (P.wide) {
        const float4* src = reinterpret_cast<const float4*>(P.x[pack_buf]) + (size_t)li * 16;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) __stcg(reinterpret_cast<float4*>(e) + k, __ldcg(src + k));
      } else {
        const double* xs = P.x[pack_buf];
        for (int k = 0; k < P.nx; ++k) __stcg(reinterpret_cast<double*>(e) + k, __ldcg(xs + (size_t)k * P.ld + li));
      }
      __stcg(reinterpret_cast<int4*>(e + P.pack_state_bytes), make_int4(gid, first, c, 0));
      pushed = 1;
      ++d;
    }
  }
  const int nlo = max(lo, own_lo), nhi = min(hi, own_hi);
  lo = nlo;
  cnt = max(nhi - nlo, 0);
}

// Cross-GPU barrier that closes the scatter of a sharded resample; it carries, per destination, the number of packed
// entries this rank pushed there.  Call right after the local grid barrier (every pushing thread executed a system-scope
// release fence before arriving at it, and block 0 posts the count with st.release.sys).
// incoming[r] = entries rank r pushed into MY pack_in (0 for r == rank).
__device__ __forceinline__ void peer_exchange_counts(const EngineP& P, Shared& sh, u64& xseq, int (&incoming)[MAX_WORLD]) {
  xseq += 1;
  const int par = (int)(xseq & 1ull);
  const u64 tag = (xseq & 0xffffffffull) << 32;
  if (LLPF_BLOCKIDX == 0 && (int)threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
    const int r = threadIdx.x;
    const unsigned c = (unsigned)__ldcg(P.pack_cnt + r);
    u64* out = reinterpret_cast<u64*>(P.peer_mbox[r]) + ((size_t)par * MAX_WORLD + P.rank) * MBOX_WORDS;
#if LLPF_XCHG_VARIANT == 1
    st_relaxed_sys_u64(out, tag | (u64)c);
#else
    st_release_sys_u64(out, tag | (u64)c);   // cumulative: everything ordered before it by the grid barrier is released with it
#endif
  }
  uint32_t* words = reinterpret_cast<uint32_t*>(sh.peer_vals);
  __syncthreads();                                               // earlier readers of sh.peer_vals are done
  if ((int)threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
    const u64* in = reinterpret_cast<const u64*>(P.peer_mbox[P.rank]) + ((size_t)par * MAX_WORLD + threadIdx.x) * MBOX_WORDS;
    u64 v;
    do { v = ld_relaxed_sys_u64(in); } while ((v & 0xffffffff00000000ull) != tag);
    words[threadIdx.x] = (uint32_t)v;
    LLPF_TS(P, sh, 15);
    // acquire side: the entries behind the count are read next (by the whole block, after the barrier below).  An acquire
    // LOAD of the word just seen, not a fence: fence.acq_rel.sys here cost 7 us per resample step on 2 GPUs (79.5 -> 72.4 us),
    // fence.sc.sys (__threadfence_system) another 1 us (profiles/r2_xchg_ab.log)
    (void)ld_acquire_sys_u64(in);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < MAX_WORLD; ++r) incoming[r] = (r < P.world && r != P.rank) ? (int)words[r] : 0;
}

// ---- FAST path, two adjacent particles per lane (16-byte loads/stores, half the shuffles per particle) ----
// Needs an even chunk start (the host rounds chunks to even sizes).  Row r covers 2*BLOCK particles.
template <class LoadFn, class WeFn>
__device__ __forceinline__ void scan_stage1_pairs(const EngineP& P, Shared& sh, int beg, int end, LoadFn loadfn, WeFn wefn) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = (end - beg + 2 * BLOCK - 1) / (2 * BLOCK);
  double r0n = 0.0, r1n = 0.0;
  {
    const int i = beg + 2 * threadIdx.x;
    if (i < end) r0n = loadfn(i);
    if (i + 1 < end) r1n = loadfn(i + 1);
  }
  for (int k = 0; k < rows; ++k) {
    const int i = beg + k * 2 * BLOCK + 2 * threadIdx.x;
    const double r0 = r0n, r1 = r1n;
    {
      const int in = i + 2 * BLOCK;
      if (in < end) r0n = loadfn(in);
      if (in + 1 < end) r1n = loadfn(in + 1);
    }
    const u64 f0 = (i < end) ? to_fixed(wefn(i, r0), P.fix_scale) : 0ull;
    const u64 f1 = (i + 1 < end) ? to_fixed(wefn(i + 1, r1), P.fix_scale) : 0ull;
    const u64 pr = f0 + f1;
    u64 v = pr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (i + 1 < end) {
      __stcg(reinterpret_cast<ulonglong2*>(P.loc + i), make_ulonglong2(v - f1, v));
    } else if (i < end) {
      __stcg(P.loc + i, v - f1);
    }
    if (lane == 31) sh.wt[k * NWARP + warp] = v;
  }
  __syncthreads();
  if (warp == 0) {   // exclusive scan of the rows*NWARP segment totals, in (row, warp) order
    const int n = rows * NWARP;
    const int per = (n + 31) >> 5;
    const int s0 = lane * per, s1 = (s0 + per < n) ? s0 + per : n;
    u64 sum = 0;
    for (int q = s0; q < s1; ++q) sum += sh.wt[q];
    u64 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    u64 run = incl - sum;
    for (int q = s0; q < s1; ++q) {
      const u64 t = sh.wt[q];
      sh.wt[q] = run;
      run += t;
    }
    if (lane == 31) __stcg(P.tots + LLPF_BLOCKIDX, incl);
  }
  __syncthreads();
}

template <class JT>
__device__ __forceinline__ void scatter_pairs(const EngineP& P, Shared& sh, int beg, int end, u64 off,
                                              const Thresholds& th, const SlotRouter<JT>& jout, JT jbase, int pack_buf,
                                              int& pushed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = (end - beg + 2 * BLOCK - 1) / (2 * BLOCK);
  // F at every (row, warp) segment start: lane 0 of a segment needs it for its first particle
  __syncthreads();
  for (int q = threadIdx.x; q < rows * NWARP; q += BLOCK)
    sh.wtf[q] = th.ipath ? first_slot_ge_fixed(th, off + sh.wt[q])
                         : first_slot_ge(th, P.key, (double)(off + sh.wt[q]) * P.fix_inv);
  __syncthreads();
  ulonglong2 mn = make_ulonglong2(0ull, 0ull);
  {
    const int i = beg + 2 * threadIdx.x;
    if (i + 1 < end) mn = __ldcg(reinterpret_cast<const ulonglong2*>(P.loc + i));
    else if (i < end) mn.x = __ldcg(P.loc + i);
  }
  for (int k = 0; k < rows; ++k) {
    const int i = beg + k * 2 * BLOCK + 2 * threadIdx.x;
    const ulonglong2 m = mn;
    {
      const int in = i + 2 * BLOCK;
      mn = make_ulonglong2(0ull, 0ull);
      if (in + 1 < end) mn = __ldcg(reinterpret_cast<const ulonglong2*>(P.loc + in));
      else if (in < end) mn.x = __ldcg(P.loc + in);
    }
    const u64 seg = off + sh.wt[k * NWARP + warp];
    int fB = 0, fC = 0;
    if (i < end) {
      const double hi0 = (double)(seg + m.x) * P.fix_inv;
      fB = th.ipath ? first_slot_ge_fixed(th, seg + m.x) : first_slot_ge(th, P.key, hi0);
      fC = fB;
      if (i + 1 < end) {
        const double hi1 = (double)(seg + m.y) * P.fix_inv;
        if (m.y > m.x) fC = th.ipath ? first_slot_ge_fixed(th, seg + m.y) : first_slot_ge(th, P.key, hi1);
        __stcg(reinterpret_cast<double2*>(P.bins + i), make_double2(hi0, hi1));
      } else {
        __stcg(P.bins + i, hi0);
      }
    }
    int fA = __shfl_up_sync(0xffffffffu, fC, 1);       // F(lo) of my first particle = F(hi) of the previous lane's second
    if (lane == 0) fA = sh.wtf[k * NWARP + warp];
    if (i >= end) { fA = 0; fB = 0; fC = 0; }
    int l0 = fA, c0 = fB - fA, l1 = fB, c1 = fC - fB;
    if (P.world > 1) {   // offspring in other ranks' slots travel as packed entries; the own part is scattered below
      push_remote_parts(P, pack_buf, l0, c0, (int)jbase + i, i, pushed);
      push_remote_parts(P, pack_buf, l1, c1, (int)jbase + i + 1, i + 1, pushed);
    }
    scatter_runs<JT>(jout, l0, c0, jbase + (JT)i);
    scatter_runs<JT>(jout, l1, c1, jbase + (JT)(i + 1));
  }
}

// After the barrier(s) that close the scatter: fill the queued heavy runs into the slots [slot_lo, slot_hi) this
// block is responsible for (the engine: exactly the slots its sweep reads next).  jout[s - slot_base] = id.
// The list is local: runs that reach into other ranks were clipped at the source and shipped as packed entries.
template <class JT>
__device__ __forceinline__ void fill_heavy_runs(const EngineP& P, JT* jout, int slot_base, int slot_lo, int slot_hi, int n) {
  if (n <= 0) return;   // block-uniform
  if (n > HEAVY_MAX) n = HEAVY_MAX;
  for (int e = 0; e < n; ++e) {
    const int lo = __ldcg(P.heavy + 1 + 3 * e), c = __ldcg(P.heavy + 2 + 3 * e);
    const JT id = (JT)__ldcg(P.heavy + 3 + 3 * e);
    const int a = max(lo, slot_lo), b = min(lo + c, slot_hi);
    for (int sl = a + threadIdx.x; sl < b; sl += BLOCK) __stcg(jout + (sl - slot_base), id);
  }
  __syncthreads();
}

// Destination side of the packed exchange: every block takes a strided share of the entries each source pushed here
// and writes the runs into the LOCAL j as negative entry codes -(1 + flat entry index); very long runs go through the
// local heavy-run list like any other.  The gathering sweep turns the codes back into global ancestor ids.
template <class JT>
__device__ __forceinline__ void expand_packs(const EngineP& P, JT* jout_flat, const int (&incoming)[MAX_WORLD]) {
  SlotRouter<JT> R;
  R.j = jout_flat; R.base = P.first; R.heavy = P.heavy;
  for (int s = 0; s < P.world; ++s) {
    const int ns = incoming[s];
    const char* meta = P.pack_in + (size_t)s * (size_t)P.pack_cap * (size_t)P.pack_stride + P.pack_state_bytes;
    for (int base = LLPF_BLOCKIDX * BLOCK; base < ns; base += P.nblocks * BLOCK) {   // block-uniform trip count
      const int idx = base + threadIdx.x;
      int first = 0, c = 0;
      if (idx < ns) {
        const int4 m = __ldcg(reinterpret_cast<const int4*>(meta + (size_t)idx * (size_t)P.pack_stride));
        first = m.y; c = m.z;
      }
      scatter_runs<JT>(R, first, c, (JT)(-1 - (s * (int)P.pack_cap + idx)));
    }
  }
}

// Closes the scatter: local grid barrier; sharded filters then exchange the packed-entry counts (the cross-GPU barrier)
// and expand what arrived; finally the heavy-run list is filled into the block's own slots.
template <class JT>
__device__ __forceinline__ void finish_scatter(const EngineP& P, Shared& sh, unsigned& bar_target, u64& xseq,
                                               JT* jout_flat, int slot_lo, int slot_hi, int pushed) {
#if LLPF_XCHG_VARIANT != 2
  if (pushed) fence_acq_rel_sys();      // release side: my packed entries are ordered before the count block 0 posts
#endif
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  LLPF_TS(P, sh, 10);
  if (P.world > 1) {
    int incoming[MAX_WORLD];
    peer_exchange_counts(P, sh, xseq, incoming);
    LLPF_TS(P, sh, 11);
    int tot = 0;
#pragma unroll
    for (int r = 0; r < MAX_WORLD; ++r) tot += incoming[r];
    if (tot > 0) {   // identical in every block of the rank
      expand_packs<JT>(P, jout_flat, incoming);
      LLPF_TS(P, sh, 12);
      grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
      LLPF_TS(P, sh, 13);
    }
  }
  const int hn = (P.heavy != nullptr) ? __ldcg(P.heavy) : 0;
  fill_heavy_runs<JT>(P, jout_flat, P.first, slot_lo, slot_hi, hn);
}

// The whole resample: scan(we) -> bins (global) -> per-source slot ranges -> j (global, id = jbase + i).
// Ends with a grid barrier: afterwards j[s] is valid for every slot s < f_total (returned); the
// remaining slots are the reference's "untouched" entries (resample.jl:26-34).
template <class JT, class LoadFn, class WeFn>
__device__ __forceinline__ int resample_indices(const EngineP& P, Shared& sh, int beg, int end, unsigned& bar_target,
                                                LoadFn loadfn, WeFn wefn, double u01, bool gen_u01, uint32_t step_idx,
                                                int Mslots, const double* u_slots, JT* jout_flat, JT jbase,
                                                double& total_out, u64& xseq, int slot_lo, int slot_hi,
                                                const RangeArg* range = nullptr, int pack_buf = 0) {
  // [slot_lo, slot_hi): the GLOBAL output slots this block fills from the heavy-run list; jout_flat[0] is global
  // slot P.first
  if (P.heavy != nullptr && LLPF_BLOCKIDX == 0 && threadIdx.x == 0) __stcg(P.heavy, 0);   // before the first barrier
  if (P.world > 1 && LLPF_BLOCKIDX == 0 && (int)threadIdx.x < P.world) __stcg(P.pack_cnt + threadIdx.x, 0);
  SlotRouter<JT> jout;
  jout.heavy = P.heavy;
  jout.j = jout_flat;
  jout.base = P.first;
  int pushed = 0;
  const int rows2 = (end - beg + 2 * BLOCK - 1) / (2 * BLOCK);
  const bool pairs = (P.scan_mode == 0) && (rows2 <= MAX_ROWS) && ((beg & 1) == 0);
  bool tabled = false;
  if (pairs) scan_stage1_pairs(P, sh, beg, end, loadfn, wefn);
  else tabled = scan_stage1(P, sh, beg, end, loadfn, wefn);
  LLPF_TS(P, sh, 1);
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  LLPF_TS(P, sh, 2);
  double total;
  u64 off = 0, gtot_fixed = 0;
  if (P.scan_mode != 0) {
    if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0) scan_serial(P.bins, P.n);
    grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
    total = __ldcg(P.bins + P.n - 1);
  } else {
    scan_block_offsets(P, sh, 0ull);
    u64 gbase = 0, gtot = sh.offs[P.nblocks];
    if (P.world > 1) {
      // global CDF offset of this rank: all-gather of the ranks' fixed-point totals (exact integers, so the
      // global bins are bit-identical to a single-GPU scan of the same weights)
      double mine[1] = {__longlong_as_double((long long)gtot)}, all[MAX_WORLD][1];
      peer_allgather<1>(P, sh, xseq, mine, all);
      gtot = 0;
      for (int r = 0; r < P.world; ++r) {
        const u64 tr = (u64)__double_as_longlong(all[r][0]);
        if (r < P.rank) gbase += tr;
        gtot += tr;
      }
    }
    total = (double)gtot * P.fix_inv;
    gtot_fixed = gtot;
    off = gbase + sh.offs[LLPF_BLOCKIDX];
    LLPF_TS(P, sh, 14);
  }
  total_out = total;
  if (gen_u01) u01 = resample_u01(P.key, step_idx);
  Thresholds th = make_thresholds(P, total, u01, Mslots, step_idx, u_slots);
  if (range != nullptr && range->rational) {   // stand-alone entry: the host found Julia's rational range for (r, 1/M, total+r)
    th.ipath = 0;
    th.twice = 1; th.offset = range->offset;
    th.r = range->ref_hi; th.ref_lo = range->ref_lo; th.step = range->step_hi; th.step_lo = range->step_lo;
  }
  if (pairs) {
    scatter_pairs<JT>(P, sh, beg, end, off, th, jout, jbase, pack_buf, pushed);
    const int f_tot = th.ipath ? first_slot_ge_fixed(th, gtot_fixed) : first_slot_ge(th, P.key, total);
    LLPF_TS(P, sh, 3);
    finish_scatter<JT>(P, sh, bar_target, xseq, jout_flat, slot_lo, slot_hi, pushed);
    LLPF_TS(P, sh, 4);
    return f_tot;
  }
  int row = 0;
  // software pipeline: the next row's prefix (FAST) / bins pair (SERIAL) is in flight while this row is processed
  u64 mine_next = 0;
  double hi_next = 0.0, lo_next = 0.0;
  {
    const int i0 = beg + threadIdx.x;
    if (i0 < end) {
      if (P.scan_mode == 0) mine_next = __ldcg(P.loc + i0);
      else { hi_next = __ldcg(P.bins + i0); lo_next = (i0 > 0) ? __ldcg(P.bins + i0 - 1) : 0.0; }
    }
  }
  for (int base = beg; base < end; base += BLOCK, ++row) {   // uniform trip count: warp collectives inside
    const int i = base + threadIdx.x;
    int f_lo = 0, cnt = 0;
    u64 mine = mine_next, seg = 0;
    const double hi_s = hi_next, lo_s = lo_next;
    {
      const int in = i + BLOCK;
      mine_next = 0;
      if (in < end) {
        if (P.scan_mode == 0) mine_next = __ldcg(P.loc + in);
        else { hi_next = __ldcg(P.bins + in); lo_next = __ldcg(P.bins + in - 1); }
      }
    }
    if (P.scan_mode == 0 && tabled) seg = sh.wt[row * NWARP + (threadIdx.x >> 5)];
    const u64 prev = __shfl_up_sync(0xffffffffu, mine, 1);   // inclusive prefix of the previous lane's particle
    if (i < end) {
      double lo, hi;
      if (P.scan_mode != 0) {
        hi = hi_s;
        lo = lo_s;
      } else {
        u64 plo;
        if (tabled) plo = ((threadIdx.x & 31) == 0) ? 0ull : prev;         // segment-relative
        else plo = (i > beg) ? (((threadIdx.x & 31) == 0) ? __ldcg(P.loc + i - 1) : prev) : 0ull;
        hi = (double)(off + seg + mine) * P.fix_inv;
        lo = (double)(off + seg + plo) * P.fix_inv;
        __stcg(P.bins + i, hi);
      }
      f_lo = first_slot_ge(th, P.key, lo);
      const int f_hi = (hi > lo) ? first_slot_ge(th, P.key, hi) : f_lo;
      cnt = f_hi - f_lo;
    }
    if (P.world > 1) push_remote_parts(P, pack_buf, f_lo, cnt, (int)jbase + i, i, pushed);
    scatter_runs<JT>(jout, f_lo, cnt, jbase + (JT)i);
  }
  const int f_total = first_slot_ge(th, P.key, total);
  LLPF_TS(P, sh, 3);
  finish_scatter<JT>(P, sh, bar_target, xseq, jout_flat, slot_lo, slot_hi, pushed);
  LLPF_TS(P, sh, 4);
  return f_total;
}

}  // namespace llpf
#include "llpf_residual.cuh"   // resample(ResampleResidual, ...) — uses the scan / scatter helpers above
#include "llpf_metropolis.cuh" // Metropolis resampling (extension; RESID == 2 instantiations)
namespace llpf {

// ------------------------------------------------------------------------------------------------
// models
// ------------------------------------------------------------------------------------------------
// quadtank right-hand side, example_quadtank.jl:91-106 (+ the t>500 leak switch of :15-17)
template <int NX, int NY>
__device__ __forceinline__ void quadtank_rhs(const ModelP<NX, NY>& M, const double (&bu)[NX], double t,
                                             const double (&h)[NX], double (&xd)[NX]) {
  const double c1 = (t > M.t_switch) ? M.qt[1] : M.qt[0];
  const double co = M.qt[0], ci = M.qt[2], tg = M.qt[3];
  const double s0 = sqrt(fmax(tg * h[0], 0.0) + 1e-3);
  const double s1 = sqrt(fmax(tg * h[1], 0.0) + 1e-3);
  const double s2 = sqrt(fmax(tg * h[2], 0.0) + 1e-3);
  const double s3 = sqrt(fmax(tg * h[3], 0.0) + 1e-3);
  xd[0] = c1 * s0 + ci * s2 + bu[0];
  xd[1] = co * s1 + ci * s3 + bu[1];
  xd[2] = co * s2 + bu[2];
  xd[3] = co * s3 + bu[3];
}

// dynamics(x,u,p,t) without noise, in place.  DYN==0: A*x .+ B*u ; DYN==1: rk4(quadtank) utils.jl:220-237
template <int NX, int NY, int DYN>
__device__ __forceinline__ void dynamics_mean(const ModelP<NX, NY>& M, const Shared& sh, const double (&bu)[NX], double t,
                                              double (&x)[NX]) {
#ifdef LLPF_USER_MODEL
  if constexpr (DYN == 2) {
    llpf_user::dynamics<NX>(x, sh.u_prop, sh.user_p, t);
  } else
#endif
  if (DYN == 0) {
    double xn[NX];
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      double a[NX];
      lds_row<NX>(sh.mA + r * mdl_stride(NX), a);
      double acc = a[0] * x[0];
#pragma unroll
      for (int c = 1; c < NX; ++c) acc = fma(a[c], x[c], acc);
      xn[r] = acc + bu[r];
    }
#pragma unroll
    for (int r = 0; r < NX; ++r) x[r] = xn[r];
  } else {
    const double h = M.integ_h;
    for (int s = 0; s < M.supersample; ++s) {
      double f1[NX], f2[NX], f3[NX], f4[NX], tmp[NX];
      quadtank_rhs<NX, NY>(M, bu, t, x, f1);
#pragma unroll
      for (int i = 0; i < NX; ++i) tmp[i] = x[i] + h / 2 * f1[i];
      quadtank_rhs<NX, NY>(M, bu, t + h / 2, tmp, f2);
#pragma unroll
      for (int i = 0; i < NX; ++i) tmp[i] = x[i] + h / 2 * f2[i];
      quadtank_rhs<NX, NY>(M, bu, t + h / 2, tmp, f3);
#pragma unroll
      for (int i = 0; i < NX; ++i) tmp[i] = x[i] + h * f3[i];
      quadtank_rhs<NX, NY>(M, bu, t + h, tmp, f4);
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] += h / 6 * (f1[i] + 2 * f2[i] + 2 * f3[i] + f4[i]);
      t += h;
    }
  }
}

// x += L1*z, z ~ N(0,I) from the (ST_DYN, step, particle) counter  (PFtypes.jl:135,153; utils.jl:260-268)
template <int NX, int NY>
__device__ __forceinline__ void add_dynamics_noise(const ModelP<NX, NY>& M, const RngKey& key,
                                                   uint32_t step_idx, int gi, double (&x)[NX], const Shared& sh) {
  double z[NX];
  normals<NX>(key, ST_DYN, step_idx, (unsigned long long)(unsigned)gi, z, sh.mt);
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    double l[NX];
    lds_row<NX>(sh.mL + r * mdl_stride(NX), l);
    double acc = x[r];
#pragma unroll
    for (int c = 0; c <= r; ++c) acc = fma(l[c], z[c], acc);
    x[r] = acc;
  }
}

// nz = L1*z (the additive dynamics noise), independent of the particle value
template <int NX, int NY>
__device__ __forceinline__ void noise_vector(const RngKey& key, uint32_t step_idx, int gi, double (&nz)[NX], const Shared& sh) {
  double z[NX];
  normals<NX>(key, ST_DYN, step_idx, (unsigned long long)(unsigned)gi, z, sh.mt);
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    double l[NX];
    lds_row<NX>(sh.mL + r * mdl_stride(NX), l);
    double acc = l[0] * z[0];
#pragma unroll
    for (int c = 1; c <= r; ++c) acc = fma(l[c], z[c], acc);
    nz[r] = acc;
  }
}

// logpdf(N(0,R2), y - C x) = c0 - |W(y - Cx)|^2/2 = c0 - |yt - G x|^2/2   (utils.jl:252-257)
template <int NX, int NY>
__device__ __forceinline__ double meas_loglik(const ModelP<NX, NY>& M, const Shared& sh, const double (&yt)[NY],
                                              const double (&x)[NX]) {
  double q = 0.0;
#pragma unroll
  for (int a = 0; a < NY; ++a) {
    double g[NX];
    lds_row<NX>(sh.mG + a * mdl_stride(NX), g);
    double v = yt[a];
#pragma unroll
    for (int c = 0; c < NX; ++c) v = fma(-g[c], x[c], v);
    q = fma(v, v, q);
  }
  return fma(-0.5, q, M.c0);
}

// w[i] += logpdf(dg, y - g(x))  (PFtypes.jl:116)  or the user's measurement_likelihood(x,u,y,p,t)  (PFtypes.jl:232)
template <int NX, int NY, int DYN>
__device__ __forceinline__ double weigh_loglik(const ModelP<NX, NY>& M, const Shared& sh, const double (&yt)[NY],
                                               const double (&x)[NX], double t) {
#ifdef LLPF_USER_MODEL
  if constexpr (DYN == 2) return llpf_user::loglik<NX>(x, sh.u_weigh, sh.y_raw, sh.user_p, t);
#endif
  return meas_loglik<NX, NY>(M, sh, yt, x);
}

template <int NX, int NY>
__device__ __forceinline__ void model_to_shared(const ModelP<NX, NY>& M, Shared& sh) {
  constexpr int S = mdl_stride(NX);
  for (int k = threadIdx.x; k < NX * S; k += BLOCK) {
    const int r = k / S, c = k % S;
    sh.mA[k] = (c < NX) ? M.A[r * NX + c] : 0.0;
    sh.mL[k] = (c < NX) ? M.L1[r * NX + c] : 0.0;
  }
  for (int k = threadIdx.x; k < NY * S; k += BLOCK) {
    const int r = k / S, c = k % S;
    sh.mG[k] = (c < NX) ? M.G[r * NX + c] : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// per-pass helpers
// ------------------------------------------------------------------------------------------------
template <int V>
struct IntTag {
  static constexpr int value = V;
};

struct Ctx {
  int beg, end;        // own chunk (local indices)
  double lwN;          // -log(N)  (filtering.jl:11)
  double lw1N;         // log(1/N) (utils.jl:75)
  unsigned bar_target;
  unsigned red_seq;    // number of reductions done in this launch (tag of the published statistics)
};

// by-value snapshot of the lazy weight state
struct WState {
  int uniform, pend;
  double pm, pls, inv_s, wu, weu;
  const double* w;
  const MathTab* T;
  __device__ __forceinline__ double weight_norm_raw(double wr) const {
    if (uniform) return wu;
    return pend ? (wr - pm) - pls : wr;
  }
  __device__ __forceinline__ double expweight_raw(double wr) const {
    if (uniform) return weu;
    return pend ? exp_nonpos(wr - pm, *T) * inv_s : exp(wr);
  }
  // logical (normalised) log-weight of local particle i: (w - offset) - log1p(s)  utils.jl:20,25
  __device__ __forceinline__ double weight_norm(int i) const {
    if (uniform) return wu;
    const double wr = __ldcg(w + i);
    return pend ? (wr - pm) - pls : wr;
  }
  // we = exp(w - offset) * 1/(s+1)   utils.jl:21-24
  __device__ __forceinline__ double expweight(int i) const {
    if (uniform) return weu;
    const double wr = __ldcg(w + i);
    return pend ? exp_nonpos(wr - pm, *T) * inv_s : exp(wr);
  }
};
__device__ __forceinline__ WState make_wstate(const EngineP& P, const Scalars& sc, const Ctx& cx, const MathTab& T) {
  WState ws;
  ws.uniform = sc.uniform; ws.pend = sc.pend;
  ws.pm = sc.pend_m; ws.pls = sc.pend_ls;
  ws.inv_s = sc.pend ? 1.0 / sc.pend_s : 1.0;
  ws.wu = (sc.uniform == 1) ? cx.lwN : cx.lw1N;
  ws.weu = 1.0 / (double)P.N;
  ws.w = P.w;
  ws.T = &T;
  return ws;
}

__device__ __forceinline__ double step_time(const EngineP& P, int k) {  // k is 1-based
  if (P.use_t_override) return P.t_override;
  return (double)(k - 1 + P.time_conv) * P.Ts;
}

// per-pass uniform data: bu = B*u_k (or the quadtank input terms), yt = W*y_k, skip = any(isnan(y))
template <int NX, int NY, int DYN>
__device__ __forceinline__ void stage_step(const EngineP& P, const ModelP<NX, NY>& M, Shared& sh, int k_u, int k_y,
                                           double (&bu)[NX], double (&yt)[NY], bool& skip) {
  __syncthreads();
#ifdef LLPF_USER_MODEL
  if constexpr (DYN == 2) {   // user-defined model: the raw vectors of the pass (u of both steps, y of the weighed one)
    if (threadIdx.x >= 64 && threadIdx.x < 64 + MAX_NU) {
      const int c = threadIdx.x - 64;
      sh.u_prop[c] = (k_u > 0 && c < M.nu) ? __ldg(P.u + (size_t)(k_u - 1) * M.nu + c) : 0.0;
      sh.u_weigh[c] = (k_y > 0 && c < M.nu) ? __ldg(P.u + (size_t)(k_y - 1) * M.nu + c) : 0.0;
      if (c < NY) sh.y_raw[c] = (k_y > 0) ? __ldg(P.y + (size_t)(k_y - 1) * NY + c) : 0.0;
    }
  }
#endif
#ifdef LLPF_USER_MODEL
  if (k_u > 0 && threadIdx.x < NX && DYN != 2) {
#else
  if (k_u > 0 && threadIdx.x < NX) {
#endif
    const double* u = P.u + (size_t)(k_u - 1) * M.nu;
    double acc = 0.0;
    if (DYN == 0) {
      for (int c = 0; c < M.nu; ++c) acc = fma(M.B[threadIdx.x * MAX_NU + c], __ldg(u + c), acc);
    } else {
      // {g1k1/A*u1, g2k2/A*u2, (1-g2)k2/A*u2, (1-g1)k1/A*u1}
      const int ui = (threadIdx.x == 0 || threadIdx.x == 3) ? 0 : 1;
      acc = M.qt[4 + threadIdx.x] * __ldg(u + ui);
    }
    sh.bu[threadIdx.x] = acc;
  }
  if (k_y > 0 && threadIdx.x == 32) {
    const double* y = P.y + (size_t)(k_y - 1) * NY;
    int sk = 0;
    double yv[NY];
#pragma unroll
    for (int a = 0; a < NY; ++a) {
      yv[a] = __ldg(y + a);
      if (isnan(yv[a])) sk = 1;
    }
#pragma unroll
    for (int a = 0; a < NY; ++a) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c <= a; ++c) acc = fma(M.W[a * NY + c], yv[c], acc);
      sh.yt[a] = acc;
    }
    sh.skip = sk;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < NX; ++r) bu[r] = (k_u > 0) ? sh.bu[r] : 0.0;
#pragma unroll
  for (int a = 0; a < NY; ++a) yt[a] = (k_y > 0) ? sh.yt[a] : 0.0;
  skip = (k_y > 0) ? (sh.skip != 0) : false;
}

// ancestor `a` of local slot `slot` from buffer `buf_id`: a >= 0 is a GLOBAL particle index (local on sharded filters
// except for the reference's "untouched" stale entries, resample.jl:26-34), a < 0 the code of a packed entry that
// another rank shipped here (expand_packs); the code is replaced by the entry's global id in state.j.
template <int NX>
__device__ __forceinline__ void gather_x(const EngineP& P, int buf_id, int a, int slot, double (&x)[NX]) {
  const double* buf;
  int li;
  if (P.world > 1) {
    if (a < 0) {
      const char* e = P.pack_in + (size_t)(-1 - a) * (size_t)P.pack_stride;
#pragma unroll
      for (int d = 0; d + 1 < NX; d += 2) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(e) + (d >> 1));
        x[d] = v.x; x[d + 1] = v.y;
      }
      if (NX & 1) x[NX - 1] = __ldcg(reinterpret_cast<const double*>(e) + (NX - 1));
      __stcg(P.j + slot, __ldcg(reinterpret_cast<const int*>(e + P.pack_state_bytes)));
      return;
    }
    li = a - P.first;
    buf = P.x[buf_id];
    if ((unsigned)li >= (unsigned)P.n) {   // a stale entry whose particle lives on another GPU (rare)
      const int r = a / P.n;
      li = a - r * P.n;
      buf = P.peer_x[r][buf_id];
    }
  } else {
    li = a - P.first;
    buf = P.x[buf_id];
  }
#pragma unroll
  for (int d = 0; d < NX; ++d) x[d] = __ldcg(buf + (size_t)d * P.ld + li);
}
template <int NX>
__device__ __forceinline__ void load_x(const double* buf, long long ld, int i, double (&x)[NX]) {
#pragma unroll
  for (int d = 0; d < NX; ++d) x[d] = __ldcg(buf + (size_t)d * ld + i);
}
template <int NX>
__device__ __forceinline__ void store_x(double* buf, long long ld, int i, const double (&x)[NX]) {
#pragma unroll
  for (int d = 0; d < NX; ++d) __stcg(buf + (size_t)d * ld + i, x[d]);
}
template <int NX>
__device__ __forceinline__ void store_hist_x(const EngineP& P, int k, int gi, const double (&x)[NX]) {
  double* p = P.x_hist + ((size_t)(k - 1) * P.N + gi) * NX;
#pragma unroll
  for (int d = 0; d < NX; ++d) __stcs(p + d, x[d]);
}

template <int NX>
__device__ __forceinline__ void publish_step(const EngineP& P, Scalars& sc, int k, const Stats& st) {
  const double ls = log(st.s);
  const double ll = st.m + ls;  // log1p(s)+offset of utils.jl:26 (s there excludes the arg-max term)
  sc.pend = 1; sc.uniform = 0; sc.stats_ahead = 0; sc.stats_valid = 1;
  sc.pend_m = st.m; sc.pend_ls = ls; sc.pend_s = st.s;
  sc.ess = st.s * st.s / st.q;   // effective_particles = 1/sum(abs2, we)  resample.jl:1-2
  sc.ll_last = ll;
  sc.ll_total += ll;
  if (!(fabs(ll) <= DBL_MAX)) sc.nonfinite = 1;
#pragma unroll
  for (int d = 0; d < NX; ++d) sc.xhat[d] = st.sx[d] / st.s;
  if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0 && k > 0) {
    if (P.ll_steps) P.ll_steps[k - 1] = ll;
    if (P.ess_steps) P.ess_steps[k - 1] = sc.ess;
    if (P.xhat) {
#pragma unroll
      for (int d = 0; d < NX; ++d) P.xhat[(size_t)(k - 1) * NX + d] = sc.xhat[d];
    }
  }
}

// ---- PF / AdvancedPF pass: [predict!(k_prop)] fused with [correct!(k_weigh)] -----------------------
// k_prop / k_weigh are 1-based step numbers, 0 = phase absent.
// RESID != 0: the kernel instantiation for resampling_strategy = ResampleResidual (a separate instantiation, so the
// systematic / stratified kernels carry none of its code: adding it as a run-time branch cost 2-4 % on resample steps)
template <int NX, int NY, int DYN, int RESID>
__device__ __forceinline__ void pf_pass(const EngineP& P, const ModelP<NX, NY>& M, Shared& sh, Scalars& sc,
                                        Ctx& cx, int k_prop, int k_weigh, int flags) {
  if (flags & OPF_RAW_WEIGHTS) { sc.pend = 0; sc.stats_ahead = 0; }
  const bool skip_meas = (flags & OPF_SKIP_MEAS) != 0;
  double bu[NX], yt[NY];
  bool nan_y;
#ifdef LLPF_PHASE_TIMING
  if (threadIdx.x == 0) sh.pass_id = (k_prop > 0) ? k_prop : 0;
#endif
  stage_step<NX, NY, DYN>(P, M, sh, k_prop, skip_meas ? 0 : k_weigh, bu, yt, nan_y);
  LLPF_TS(P, sh, 0);
  LLPF_TSB(P, sh, P.dbg_T, 0);
  const bool skip = skip_meas || nan_y;
  // shouldresample(pf)  resample.jl:5-10
  const bool res = (k_prop > 0) && ((P.thr == 1.0) || (sc.ess < (double)P.N * P.thr));
  const bool hist_w = (P.w_hist != nullptr) && (k_prop > 0);   // weights of step k_prop, just corrected
  const WState ws = make_wstate(P, sc, cx, sh.mt);
  const uint32_t step_idx = (uint32_t)sc.t_index;
  const size_t hbase = hist_w ? (size_t)(k_prop - 1) * P.N + P.first : 0;
  double* const wh = P.w_hist;
  double* const weh = P.we_hist;
  int f_total = 0;
  if constexpr (RESID == 2) {
    if (res) {   // Metropolis resampling (llpf_metropolis.cuh; not in the reference)
      resample_metropolis(P, sh, cx.beg, cx.end, cx.bar_target, ws.uniform ? nullptr : ws.w, step_idx,
                          [=](int i, double wr) {
                            if (hist_w) {
                              __stcs(wh + hbase + i, ws.weight_norm_raw(wr));
                              __stcs(weh + hbase + i, ws.expweight_raw(wr));
                            }
                          });
      f_total = (int)P.N;
      sc.bins_total = 1.0;
    }
  } else if constexpr (RESID != 0) {
    if (res) {   // ResampleResidual  resample.jl:63-117
      WeSrc rs;
      rs.w = ws.w; rs.mode = ws.uniform ? 1 : (ws.pend ? 3 : 2);
      rs.pm = ws.pm; rs.pls = ws.pls; rs.inv_s = ws.inv_s; rs.weu = ws.weu; rs.wu = ws.wu; rs.T = &sh.mt;
      rs.hist_w = hist_w ? wh + hbase : nullptr;
      rs.hist_we = hist_w ? weh + hbase : nullptr;
      double total;
      resample_residual<int>(P, sh, cx.beg, cx.end, cx.bar_target, rs, nullptr, step_idx, (int)P.N, P.j, P.first,
                             sc.j_identity, P.first + cx.beg, P.first + cx.end, total);
      f_total = (int)P.N;
      sc.bins_total = total;
    }
  } else if (res) {
    double total;
    f_total = resample_indices<int>(
        P, sh, cx.beg, cx.end, cx.bar_target,
        [=](int i) { return ws.uniform ? 0.0 : __ldcg(ws.w + i); },
        [=](int i, double wr) {
          const double we = ws.expweight_raw(wr);
          if (hist_w) {
            __stcs(wh + hbase + i, ws.weight_norm_raw(wr));
            __stcs(weh + hbase + i, we);
          }
          return we;
        },
        0.0, true, step_idx, (int)P.N, nullptr, P.j, P.first, total, sc.xseq, P.first + cx.beg, P.first + cx.end,
        nullptr, sc.cur);
    sc.bins_total = total;
    LLPF_TSB(P, sh, P.dbg_T, 3);
  }
  const double* src = P.x[sc.cur];
  double* dst = res ? P.x[sc.cur ^ 1] : P.x[sc.cur];
  const double tprop = step_time(P, k_prop);
  const bool with_x = (P.want_xhat != 0);
  const int jid = sc.j_identity;
  Online<NX> acc;
  acc.init();
  // The sweep is instantiated three times: MODE 1 / MODE 2 are the steady-state passes of a trajectory
  // (predict!(k) fused with correct!(k+1), no history, no missing measurement, no running mean) without /
  // with resampling, in which every loop-invariant condition is a compile-time constant; MODE 0 is the
  // general loop (first/last pass, step verbs, history, NaN measurements, xhat).
  // software pipeline: the ancestor index of the NEXT iteration is fetched one iteration ahead; within an
  // iteration the particle/weight loads are issued first, the (data-independent) noise is computed while
  // they are in flight, and only then are they consumed.
  // resample path: ancestor index fetched TWO iterations ahead, the gathered particle ONE iteration ahead
  // (it may live in a peer GPU's memory: ~2 us over NVLink)
  auto sweep = [&](auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;
    const bool res_ = (MODE == 2) ? true : (MODE == 1) ? false : res;
    const bool prop_ = (MODE != 0) ? true : (k_prop > 0);
    const bool weigh_ = (MODE != 0) ? true : (k_weigh > 0);
    const bool hist_w_ = (MODE != 0) ? false : hist_w;
    const bool hist_x_ = (MODE != 0) ? false : (P.x_hist != nullptr);
    const bool skip_ = (MODE != 0) ? false : skip;
    const bool with_x_ = (MODE != 0) ? false : with_x;
    int a_n1 = 0, a_n2 = 0;
    double xn[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) xn[d] = 0.0;
    if (res_) {
      const int i0 = cx.beg + threadIdx.x;
      if (i0 < cx.end) {
        a_n1 = __ldcg(P.j + i0);
        if (P.first + i0 >= f_total) {
          if (jid) a_n1 = P.first + i0;
          __stcg(P.j + i0, a_n1);
        }
        gather_x<NX>(P, sc.cur, a_n1, i0, xn);
      }
      if (i0 + BLOCK < cx.end) a_n2 = __ldcg(P.j + i0 + BLOCK);
    }
    for (int i = cx.beg + threadIdx.x; i < cx.end; i += BLOCK) {
      const int gi = P.first + i;
      double x[NX];
      double wraw = 0.0;
      if (res_) {
#pragma unroll
        for (int d = 0; d < NX; ++d) x[d] = xn[d];
        const int in = i + BLOCK;
        if (in < cx.end) {
          int a = a_n2;
          if (P.first + in >= f_total) {     // untouched entry (resample.jl:26-34): keep state.j
            if (jid) a = P.first + in;
            __stcg(P.j + in, a);
          }
          gather_x<NX>(P, sc.cur, a, in, xn);
          if (in + BLOCK < cx.end) a_n2 = __ldcg(P.j + in + BLOCK);
        }
      } else {
        load_x<NX>(src, P.ld, i, x);
        if (!ws.uniform) wraw = __ldcg(P.w + i);
      }
      double z[NX];
      if (prop_) noise_vector<NX, NY>(P.key, step_idx, gi, z, sh);
      double wv;
      if (res_) {
        wv = cx.lw1N;                    // reset_weights!  utils.jl:75
      } else {
        wv = ws.uniform ? ws.wu : (ws.pend ? (wraw - ws.pm) - ws.pls : wraw);
        if (hist_w_) {
          __stcs(wh + hbase + i, wv);
          __stcs(weh + hbase + i, ws.expweight(i));
        }
      }
#if defined(LLPF_USER_MODEL) && defined(LLPF_USER_STATE_HOOKS)
      // particles with per-particle sufficient statistics (llpf_user::add_noise / correct_state above): the particle is
      // stored after correct! has mutated it, also on a pass that only weighs
      static_assert(DYN == 2, "state hooks belong to user-defined models");
      if (prop_) {
        double xp[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) xp[d] = x[d];
        dynamics_mean<NX, NY, DYN>(M, sh, bu, tprop, x);
        llpf_user::add_noise<NX>(x, xp, z, sh.u_prop, sh.user_p, tprop);
      }
      if (weigh_) {
        if (!skip_) {
          const double tw = step_time(P, k_weigh);
          wv += weigh_loglik<NX, NY, DYN>(M, sh, yt, x, tw);
          llpf_user::correct_state<NX>(x, sh.u_weigh, sh.y_raw, sh.user_p, tw);
        }
        if (hist_x_) store_hist_x<NX>(P, k_weigh, gi, x);
        __stcg(P.w + i, wv);
        acc.add(wv, x, with_x_, sh.mt);
      }
      if (prop_ || (weigh_ && !skip_)) store_x<NX>(dst, P.ld, i, x);
#else
      if (prop_) {
        dynamics_mean<NX, NY, DYN>(M, sh, bu, tprop, x);
#pragma unroll
        for (int d = 0; d < NX; ++d) x[d] += z[d];
        store_x<NX>(dst, P.ld, i, x);
      }
      if (weigh_) {
        if (hist_x_) store_hist_x<NX>(P, k_weigh, gi, x);
#ifdef LLPF_USER_MODEL
        if (!skip_) wv += weigh_loglik<NX, NY, DYN>(M, sh, yt, x, step_time(P, k_weigh));
#else
        if (!skip_) wv += meas_loglik<NX, NY>(M, sh, yt, x);
#endif
        __stcg(P.w + i, wv);
        acc.add(wv, x, with_x_, sh.mt);
      }
#endif
    }
  };
  const bool steady = (k_prop > 0) && (k_weigh > 0) && !hist_w && (P.x_hist == nullptr) && !skip && !with_x;
  if (steady && res) sweep(IntTag<2>{});
  else if (steady) sweep(IntTag<1>{});
  else sweep(IntTag<0>{});
  LLPF_TS(P, sh, 5);
  LLPF_TSB(P, sh, P.dbg_T, 1);
  if (k_prop > 0) {
    if (res) {
      sc.cur ^= 1;
      sc.uniform = 2; sc.pend = 0; sc.stats_ahead = 0;
      sc.ess = (double)P.N; sc.stats_valid = 1;
      sc.j_identity = 0;
      sc.resample_count += 1;
    } else {
      sc.j_identity = 1;   // s.j .= 1:N  filtering.jl:148
    }
    sc.last_resampled = res ? 1 : 0;
    if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0 && P.resampled) P.resampled[k_prop - 1] = res ? 1 : 0;
    sc.t_index += 1;       // filtering.jl:152
  }
  if (k_weigh > 0) {
    const Stats st = reduce_stats<NX>(P, sh, acc, with_x, cx.red_seq, sc.xseq);
    LLPF_TSB(P, sh, P.dbg_T, 2);
    publish_step<NX>(P, sc, k_weigh, st);
  }
}

// ---- AuxiliaryParticleFilter predict!(pfa,u,y1,p,t)  filtering.jl:195-217 (and :219-234) -------------
// A: xbar = f(x) (no noise) ; lam = logpdf(y1 - C xbar) ; v = w + lam ; expnormalize!(v)  -> W
// scan(W) -> bins -> j
// B: x = xbar[j] + L z ; w = lam - log N (UNresampled lam, :210-213) ; stats of the new w == next correct!
template <int NX, int NY, int DYN, int RESID>
__device__ __forceinline__ void aux_step(const EngineP& P, const ModelP<NX, NY>& M, Shared& sh, Scalars& sc,
                                         Ctx& cx, int k, int k_y1) {
  double bu[NX], yt[NY];
  bool skip;
  stage_step<NX, NY, DYN>(P, M, sh, k, k_y1, bu, yt, skip);
#ifdef LLPF_USER_MODEL
  if constexpr (DYN == 2) {   // measurement_equation!(pf, u, y1, p, t, λ) gets the u of THIS step (filtering.jl:202)
    if (threadIdx.x < MAX_NU) sh.u_weigh[threadIdx.x] = sh.u_prop[threadIdx.x];
    __syncthreads();
  }
#endif
  const bool adv = (P.filter == 3);
  const double tprop = step_time(P, k);
  const double* cur = P.x[sc.cur];
  double* oth = P.x[sc.cur ^ 1];
  const bool hist_w = (P.w_hist != nullptr);
  const WState ws = make_wstate(P, sc, cx, sh.mt);
  const uint32_t step_idx = (uint32_t)sc.t_index;
  const size_t hbase = hist_w ? (size_t)(k - 1) * P.N + P.first : 0;
  Online<NX> acc;
  acc.init();
  for (int i = cx.beg + threadIdx.x; i < cx.end; i += BLOCK) {
    double x[NX];
    load_x<NX>(cur, P.ld, i, x);
    const double wn = ws.weight_norm(i);
    if (hist_w) {
      __stcs(P.w_hist + hbase + i, wn);
      __stcs(P.we_hist + hbase + i, ws.expweight(i));
    }
    dynamics_mean<NX, NY, DYN>(M, sh, bu, tprop, x);                 // :199 no noise
    if (!adv) store_x<NX>(oth, P.ld, i, x);
#ifdef LLPF_USER_MODEL
    const double lam = skip ? 0.0 : weigh_loglik<NX, NY, DYN>(M, sh, yt, x, tprop);   // same t as the propagation (:202)
#else
    const double lam = skip ? 0.0 : meas_loglik<NX, NY>(M, sh, yt, x);  // :200-202
#endif
    __stcg(P.lam + i, lam);
    const double v = wn + lam;                                    // :203
    __stcg(P.w + i, v);
    acc.add(v, x, false, sh.mt);
  }
  const Stats s1 = reduce_stats<NX>(P, sh, acc, false, cx.red_seq, sc.xseq);
  const double inv1 = 1.0 / s1.s;
  const double m1 = s1.m;
  const double* wraw = P.w;
  const MathTab* mtp = &sh.mt;
  // expnormalize!(w): exp(w-offset)*1/(s+1)   utils.jl:57-63 ; then resample (always)  :205
  double total;
  int f_total;
  if constexpr (RESID == 2) {   // Metropolis resampling on v = w + lambda (ratios: normalisation not needed)
    resample_metropolis(P, sh, cx.beg, cx.end, cx.bar_target, wraw, step_idx, [](int, double) {});
    f_total = (int)P.N;
    total = 1.0;
  } else if constexpr (RESID != 0) {   // ResampleResidual
    WeSrc rs;
    rs.w = wraw; rs.mode = 3;
    rs.pm = m1; rs.pls = 0.0; rs.inv_s = inv1; rs.weu = 0.0; rs.wu = 0.0; rs.T = mtp;
    rs.hist_w = nullptr; rs.hist_we = nullptr;
    resample_residual<int>(P, sh, cx.beg, cx.end, cx.bar_target, rs, nullptr, step_idx, (int)P.N, P.j, P.first,
                           sc.j_identity, P.first + cx.beg, P.first + cx.end, total);
    f_total = (int)P.N;
  } else {
    f_total = resample_indices<int>(
        P, sh, cx.beg, cx.end, cx.bar_target, [=](int i) { return __ldcg(wraw + i); },
        [=](int, double wr) { return exp_nonpos(wr - m1, *mtp) * inv1; }, 0.0, true,
        step_idx, (int)P.N, nullptr, P.j, P.first, total, sc.xseq, P.first + cx.beg, P.first + cx.end,
        nullptr, adv ? sc.cur : (sc.cur ^ 1));   // sweep B gathers xbar (other buffer) / AdvancedPF: xprev
  }
  sc.bins_total = total;
  const bool with_x = (P.want_xhat != 0);
  const double lN = log((double)P.N);
  const int jid = sc.j_identity;
  acc.init();
  for (int i = cx.beg + threadIdx.x; i < cx.end; i += BLOCK) {
    const int gi = P.first + i;
    int a;
    if (gi < f_total) {
      a = __ldcg(P.j + i);
    } else {
      a = jid ? gi : __ldcg(P.j + i);
      __stcg(P.j + i, a);
    }
    double x[NX];
    double wnew;
    if (adv) {
      gather_x<NX>(P, sc.cur, a, i, x);                          // :230 propagate again from xprev[j]
      dynamics_mean<NX, NY, DYN>(M, sh, bu, tprop, x);
      add_dynamics_noise<NX, NY>(M, P.key, step_idx, gi, x, sh);
      store_x<NX>(oth, P.ld, i, x);
      wnew = cx.lw1N;                                            // :228 reset_weights!
    } else {
      gather_x<NX>(P, sc.cur ^ 1, a, i, x);                      // :207 permute_with_buffer!
      add_dynamics_noise<NX, NY>(M, P.key, step_idx, gi, x, sh);     // :208 add_noise!
      store_x<NX>(P.x[sc.cur], P.ld, i, x);
      wnew = __ldcg(P.lam + i) - lN;                             // :210-213
    }
    if (P.x_hist) store_hist_x<NX>(P, k + 1, gi, x);
    __stcg(P.w + i, wnew);
    acc.add(wnew, x, with_x, sh.mt);
  }
  if (adv) sc.cur ^= 1;
  sc.j_identity = 0;
  sc.resample_count += 1;
  sc.last_resampled = 1;
  if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0 && P.resampled) P.resampled[k - 1] = 1;
  sc.t_index += 1;                                               // :215
  const Stats s2 = reduce_stats<NX>(P, sh, acc, with_x, cx.red_seq, sc.xseq);
  // stats of the raw w[] are ready; correct! (filtering.jl:170-174) has not been *called* yet
  sc.pend = 0; sc.uniform = 0;
  sc.stats_ahead = 1; sc.stats_valid = 0;
  sc.pend_m = s2.m; sc.pend_ls = log(s2.s); sc.pend_s = s2.s;
  sc.ess = s2.s * s2.s / s2.q;
#pragma unroll
  for (int d = 0; d < NX; ++d) sc.xhat[d] = s2.sx[d] / s2.s;
}

// correct!(pfa,...) = logsumexp!(state) only  filtering.jl:170-174, using the stats found by aux_step
template <int NX>
__device__ __forceinline__ void aux_correct_from_stats(const EngineP& P, Scalars& sc, int k) {
  const double ll = sc.pend_m + sc.pend_ls;
  sc.pend = 1; sc.stats_ahead = 0; sc.stats_valid = 1;
  sc.ll_last = ll;
  sc.ll_total += ll;
  if (!(fabs(ll) <= DBL_MAX)) sc.nonfinite = 1;
  if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0 && k > 0) {
    if (P.ll_steps) P.ll_steps[k - 1] = ll;
    if (P.ess_steps) P.ess_steps[k - 1] = sc.ess;
    if (P.xhat) {
#pragma unroll
      for (int d = 0; d < NX; ++d) P.xhat[(size_t)(k - 1) * NX + d] = sc.xhat[d];
    }
  }
}

// write the weight history of step k for filters whose last step has no following pass
__device__ __forceinline__ void flush_weight_history(const EngineP& P, Shared& sh, const Scalars& sc, const Ctx& cx, int k) {
  if (!P.w_hist) return;
  const WState ws = make_wstate(P, sc, cx, sh.mt);
  const size_t hbase = (size_t)(k - 1) * P.N + P.first;
  for (int i = cx.beg + threadIdx.x; i < cx.end; i += BLOCK) {
    __stcs(P.w_hist + hbase + i, ws.weight_norm(i));
    __stcs(P.we_hist + hbase + i, ws.expweight(i));
  }
}

// the body of the persistent kernel: executes the op list on this block's chunk of the particles
template <int NX, int NY, int DYN, int RESID>
__device__ __forceinline__ void engine_body(const EngineP& P, const ModelP<NX, NY>& M, Shared& sh) {
  math_tab_load(sh.mt);
  model_to_shared<NX, NY>(M, sh);
#ifdef LLPF_USER_MODEL
  if constexpr (DYN == 2) {
    if (threadIdx.x == 0) sh.user_p = P.user_p;
  }
#endif
  __syncthreads();
  Scalars sc = *P.sc;   // every block carries an identical copy in registers; block 0 writes it back
  Ctx cx;
  cx.bar_target = 0;
  cx.red_seq = 0;
  {
    long long b = (long long)LLPF_BLOCKIDX * P.chunk;
    long long e = b + P.chunk;
    if (b > P.n) b = P.n;
    if (e > P.n) e = P.n;
    cx.beg = (int)b; cx.end = (int)e;
    asm volatile("" : "+r"(cx.beg), "+r"(cx.end));   // keep the bounds in registers (no per-iteration re-derivation)
  }
#ifdef LLPF_PHASE_TIMING
  if (P.dbg && threadIdx.x == 0) {   // which SM runs this block (row of pass 0, slot 3)
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    P.dbg[(size_t)16 * (P.dbg_T + 2) + (size_t)blockIdx.x * 4 + 3] = (long long)smid;
  }
#endif
  cx.lwN = -log((double)P.N);
  cx.lw1N = log(1.0 / (double)P.N);
  for (int r = 0; r < P.nops; ++r) {
    const int kind = P.ops[r].kind, a0 = P.ops[r].a0, b0 = P.ops[r].b0, count = P.ops[r].count;
    const int da = P.ops[r].da, db = P.ops[r].db, flags = P.ops[r].flags;
    for (int c = 0; c < count; ++c) {
      int a = a0 + c * da, b = b0 + c * db, fl = flags;
      asm volatile("" : "+r"(a), "+r"(b), "+r"(fl));   // keep the decoded op in registers (no re-decode per particle)
      if (kind == OP_PF) {
        pf_pass<NX, NY, DYN, RESID>(P, M, sh, sc, cx, a, b, fl);
      } else if (kind == OP_AUX_STEP) {
        aux_step<NX, NY, DYN, RESID>(P, M, sh, sc, cx, a, b);
        if (fl & OPF_POST_CSTATS) aux_correct_from_stats<NX>(P, sc, a + 1);
      } else if (kind == OP_AUX_CSTATS) {
        aux_correct_from_stats<NX>(P, sc, a);
      } else {
        flush_weight_history(P, sh, sc, cx, a);
      }
    }
  }
  // every block must have read the incoming scalars before block 0 overwrites them
  grid_barrier(P.bar, (unsigned)P.nblocks, cx.bar_target);
  if (LLPF_BLOCKIDX == 0 && threadIdx.x == 0) *P.sc = sc;
}

// one entry of the batched multi-chain launch (llpf_engine_batch.cu): everything one block needs to run its filter
template <int NX, int NY>
struct BatchItem {
  EngineP P;
  ModelP<NX, NY> M;
  double mu0[MAX_NX];
  double L0[MAX_NX * MAX_NX];   // row-major lower Cholesky factor of Sigma0
  Scalars sc0;                  // state after reset!
};

#ifndef LLPF_BATCH
template <int NX, int NY, int DYN, int RESID>
__global__ void __launch_bounds__(BLOCK, LLPF_MIN_BLOCKS)
k_engine(const __grid_constant__ EngineP P, const __grid_constant__ ModelP<NX, NY> M) {
  __shared__ Shared sh;
  engine_body<NX, NY, DYN, RESID>(P, M, sh);
}
#else
// ---- batched multi-chain engine (PMMH calls loglik thousands of times on small filters) ------------
// One thread block = one independent filter (its own handle: arena, model, seed, epoch), created with
// llpf_config.single_block = 1 so that the handle's own launches use the same one-block geometry: a chain evaluated
// here is bit-identical to the same chain run alone.  The block performs reset! (initial particles, filtering.jl:4-14;
// same counters and arithmetic as k_init) and then the op list, with block-level synchronisation only.
template <int NX, int NY, int DYN, int RESID>
__global__ void __launch_bounds__(BLOCK, LLPF_MIN_BLOCKS)
k_engine_batch(const BatchItem<NX, NY>* __restrict__ items, Scalars* __restrict__ sc_out) {
  __shared__ Shared sh;
  __shared__ BatchItem<NX, NY> it;
  static_assert(sizeof(BatchItem<NX, NY>) % 8 == 0, "copied by 8-byte words");
  {
    const u64* src = reinterpret_cast<const u64*>(items + blockIdx.x);
    u64* dst = reinterpret_cast<u64*>(&it);
    for (int k = threadIdx.x; k < (int)(sizeof(BatchItem<NX, NY>) / 8); k += BLOCK) dst[k] = src[k];
  }
  math_tab_load(sh.mt);
  __syncthreads();
  const EngineP& P = it.P;
  for (int k = threadIdx.x; k < BAR_TOTAL_WORDS; k += BLOCK) __stcg(P.bar + k, 0u);
  // reset!(pf): xprev[i] = rand(rng, initial_density) = mu0 + L0 z   (filtering.jl:8, utils.jl:260)
  for (int i = threadIdx.x; i < P.n; i += BLOCK) {
    double z[NX];
    normals<NX>(P.key, ST_INIT, 0u, (unsigned long long)(P.first + i), z, sh.mt);
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c <= r; ++c) acc = fma(it.L0[r * MAX_NX + c], z[c], acc);
      __stcg(P.x[0] + (size_t)r * P.ld + i, it.mu0[r] + acc);
    }
  }
  if (threadIdx.x == 0) *P.sc = it.sc0;
  __threadfence();
  __syncthreads();
  engine_body<NX, NY, DYN, RESID>(P, it.M, sh);
  if (threadIdx.x == 0) sc_out[blockIdx.x] = *P.sc;
}
#endif

}  // namespace llpf
