// llpf_api.cu — host side of libllpf_b200.so: the C-ABI of include/llpf.h on top of the persistent
// sm_100a engine (llpf_engine.cuh).  No torch types, no CPU fallback: without a CUDA device every
// entry point fails with LLPF_ERR_NO_DEVICE / LLPF_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/llpf.h"
#include "llpf_engine.cuh"
#include "llpf_engine_list.h"
#include "llpf_julia_range.h"
#include "llpf_smooth.cuh"
#include "llpf_stats.cuh"
#include "llpf_wide_common.cuh"

using namespace llpf;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(expr)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      const int code_ = (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver)        \
                            ? LLPF_ERR_NO_DEVICE : LLPF_ERR_CUDA;                             \
      return fail(code_, std::string(#expr) + ": " + cudaGetErrorString(e_));                 \
    }                                                                                         \
  } while (0)
#define OKR(expr)              \
  do {                         \
    const int rc_ = (expr);    \
    if (rc_ != LLPF_OK) return rc_; \
  } while (0)

extern "C" const char* llpf_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------------
// small host linear algebra (column-major in, as the C-ABI delivers it)
// ------------------------------------------------------------------------------------------------
#define CMH(M, r, c, ld) ((M)[(size_t)(c) * (ld) + (r)])
static bool chol_lower(const double* S, int n, std::vector<double>& L) {
  L.assign((size_t)n * n, 0.0);
  for (int jc = 0; jc < n; ++jc) {
    double d = CMH(S, jc, jc, n);
    for (int k = 0; k < jc; ++k) d -= CMH(L, jc, k, n) * CMH(L, jc, k, n);
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    CMH(L, jc, jc, n) = d;
    for (int i = jc + 1; i < n; ++i) {
      double v = CMH(S, i, jc, n);
      for (int k = 0; k < jc; ++k) v -= CMH(L, i, k, n) * CMH(L, jc, k, n);
      CMH(L, i, jc, n) = v / d;
    }
  }
  return true;
}
// Positive SEMI-definite variant for user-defined models whose particles carry deterministic components (the mean and
// covariance of a per-particle Kalman filter, rbpf.jl:136-150: only the nonlinear sub-state is random): a zero pivot is
// accepted when the rest of its column is zero too, and gives a zero column of L.
static bool chol_lower_psd(const double* S, int n, std::vector<double>& L) {
  L.assign((size_t)n * n, 0.0);
  for (int jc = 0; jc < n; ++jc) {
    double d = CMH(S, jc, jc, n);
    for (int k = 0; k < jc; ++k) d -= CMH(L, jc, k, n) * CMH(L, jc, k, n);
    if (!(d >= 0.0)) return false;
    if (d == 0.0) {
      for (int i = jc + 1; i < n; ++i) {
        double v = CMH(S, i, jc, n);
        for (int k = 0; k < jc; ++k) v -= CMH(L, i, k, n) * CMH(L, jc, k, n);
        if (v != 0.0) return false;
      }
      continue;
    }
    d = std::sqrt(d);
    CMH(L, jc, jc, n) = d;
    for (int i = jc + 1; i < n; ++i) {
      double v = CMH(S, i, jc, n);
      for (int k = 0; k < jc; ++k) v -= CMH(L, i, k, n) * CMH(L, jc, k, n);
      CMH(L, i, jc, n) = v / d;
    }
  }
  return true;
}
// inverse of a lower-triangular matrix (column-major)
static void inv_lower(const std::vector<double>& L, int n, std::vector<double>& W) {
  W.assign((size_t)n * n, 0.0);
  for (int c = 0; c < n; ++c) {
    CMH(W, c, c, n) = 1.0 / CMH(L, c, c, n);
    for (int r = c + 1; r < n; ++r) {
      double acc = 0.0;
      for (int k = c; k < r; ++k) acc += CMH(L, r, k, n) * CMH(W, k, c, n);
      CMH(W, r, c, n) = -acc / CMH(L, r, r, n);
    }
  }
}

struct HostModel {
  int nx = 0, nu = 0, ny = 0, dyn = 0;
  bool wide = false;   // Float32 particles, nx/ny up to 64: the engine of llpf_wide.cuh
  std::vector<double> A, B, C, mu0, L0, L1, L2, W, G, R2;  // column-major
  double c0 = 0;
  double dynp[8] = {0}, t_switch = 0, a1_factor = 1, integ_Ts = 1;
  int supersample = 1;
};

static int build_host_model(const llpf_model* m, HostModel& H, bool wide) {
  if (!m) return fail(LLPF_ERR_BAD_ARG, "model is null");
  const int nx = m->nx, nu = m->nu, ny = m->ny;
  H.wide = wide;
  if (wide) {
    if (nx < 1 || nx > WNX || ny < 1 || ny > WNX || nu < 0 || nu > MAX_NU)
      return fail(LLPF_ERR_UNSUPPORTED, "Float32 particles: supported dimensions 1<=nx<=64, 1<=ny<=64, 0<=nu<=8");
    if (m->dynamics != LLPF_DYN_LINEAR) return fail(LLPF_ERR_UNSUPPORTED, "Float32 particles: linear dynamics only");
  } else if (nx < 1 || nx > MAX_NX || ny < 1 || ny > 8 || nu < 0 || nu > MAX_NU)
    return fail(LLPF_ERR_UNSUPPORTED, "supported dimensions: 1<=nx<=8, 1<=ny<=8, 0<=nu<=8 (Float64 particles); "
                                      "up to 64 states with particle_dtype = LLPF_PARTICLE_F32");
  H.nx = nx; H.nu = nu; H.ny = ny; H.dyn = m->dynamics;
  if (m->dynamics == LLPF_DYN_USER) {
    // dynamics mean and measurement log-likelihood are device functions (llpf_create_user): only the noise and the
    // initial density come from the descriptor; the whitened measurement matrices of the Gaussian case stay neutral
    if (wide) return fail(LLPF_ERR_UNSUPPORTED, "user-defined models: Float64 particles only");
    if (!m->R1 || !m->mu0 || !m->Sigma0) return fail(LLPF_ERR_BAD_ARG, "user-defined model: R1, mu0, Sigma0 are required");
    H.A.clear(); H.B.clear();
    H.C.assign((size_t)ny * nx, 0.0);
    H.mu0.assign(m->mu0, m->mu0 + nx);
    if (!chol_lower_psd(m->R1, nx, H.L1)) return fail(LLPF_ERR_NOT_POSDEF, "R1 is not positive semi-definite");
    if (!chol_lower_psd(m->Sigma0, nx, H.L0)) return fail(LLPF_ERR_NOT_POSDEF, "Sigma0 is not positive semi-definite");
    H.L2.assign((size_t)ny * ny, 0.0); H.W.assign((size_t)ny * ny, 0.0);
    for (int i = 0; i < ny; ++i) { CMH(H.L2, i, i, ny) = 1.0; CMH(H.W, i, i, ny) = 1.0; }
    H.G.assign((size_t)ny * nx, 0.0);
    H.c0 = 0.0;
    std::memcpy(H.dynp, m->dyn_params, sizeof(H.dynp));
    H.t_switch = m->t_switch; H.a1_factor = m->a1_factor; H.integ_Ts = m->integ_Ts;
    H.supersample = m->supersample > 0 ? m->supersample : 1;
    return LLPF_OK;
  }
  if (!m->C || !m->R1 || !m->R2 || !m->mu0 || !m->Sigma0) return fail(LLPF_ERR_BAD_ARG, "null model matrix");
  if (m->dynamics == LLPF_DYN_LINEAR) {
    if (!m->A || (nu > 0 && !m->B)) return fail(LLPF_ERR_BAD_ARG, "linear dynamics need A (and B)");
    H.A.assign(m->A, m->A + (size_t)nx * nx);
    if (nu > 0) H.B.assign(m->B, m->B + (size_t)nx * nu); else H.B.clear();
  } else if (m->dynamics == LLPF_DYN_QUADTANK_RK4) {
    if (nx != 4 || nu != 2) return fail(LLPF_ERR_BAD_ARG, "quadtank needs nx=4, nu=2");
    if (m->supersample < 1) return fail(LLPF_ERR_BAD_ARG, "supersample must be positive");
  } else {
    return fail(LLPF_ERR_BAD_ARG, "unknown dynamics kind");
  }
  H.C.assign(m->C, m->C + (size_t)ny * nx);
  H.R2.assign(m->R2, m->R2 + (size_t)ny * ny);
  H.mu0.assign(m->mu0, m->mu0 + nx);
  if (!chol_lower(m->R1, nx, H.L1)) return fail(LLPF_ERR_NOT_POSDEF, "R1 is not positive definite");
  if (!chol_lower(m->R2, ny, H.L2)) return fail(LLPF_ERR_NOT_POSDEF, "R2 is not positive definite");
  if (!chol_lower(m->Sigma0, nx, H.L0)) return fail(LLPF_ERR_NOT_POSDEF, "Sigma0 is not positive definite");
  inv_lower(H.L2, ny, H.W);
  H.G.assign((size_t)ny * nx, 0.0);
  for (int a = 0; a < ny; ++a)
    for (int c = 0; c < nx; ++c) {
      double acc = 0.0;
      for (int k = 0; k <= a; ++k) acc += CMH(H.W, a, k, ny) * CMH(H.C, k, c, ny);
      CMH(H.G, a, c, ny) = acc;
    }
  double ld = 0.0;
  for (int i = 0; i < ny; ++i) ld += std::log(CMH(H.L2, i, i, ny));
  ld *= 2;
  H.c0 = -((double)ny * std::log(2 * M_PI) + ld) / 2;  // mvnormal_c0, utils.jl:254-257
  std::memcpy(H.dynp, m->dyn_params, sizeof(H.dynp));
  H.t_switch = m->t_switch; H.a1_factor = m->a1_factor; H.integ_Ts = m->integ_Ts;
  H.supersample = m->supersample;
  return LLPF_OK;
}

template <int NX, int NY>
static void fill_modelp(const HostModel& H, ModelP<NX, NY>& M) {
  std::memset(&M, 0, sizeof(M));
  for (int r = 0; r < NX; ++r)
    for (int c = 0; c < NX; ++c) {
      if (!H.A.empty()) M.A[r * NX + c] = CMH(H.A, r, c, NX);
      M.L1[r * NX + c] = CMH(H.L1, r, c, NX);
    }
  for (int a = 0; a < NY; ++a) {
    for (int c = 0; c < NX; ++c) M.G[a * NX + c] = CMH(H.G, a, c, NY);
    for (int c = 0; c < NY; ++c) M.W[a * NY + c] = CMH(H.W, a, c, NY);
  }
  for (int r = 0; r < NX; ++r)
    for (int c = 0; c < H.nu; ++c)
      if (!H.B.empty()) M.B[r * MAX_NU + c] = CMH(H.B, r, c, NX);
  M.c0 = H.c0;
  if (H.dyn == LLPF_DYN_QUADTANK_RK4) {
    // p = {kc,k1,k2,A,a,gamma}  example_quadtank.jl:91-97 ; coefficients in the reference's evaluation order
    const double k1 = H.dynp[1], k2 = H.dynp[2], Aa = H.dynp[3], a = H.dynp[4], g = H.dynp[5];
    M.qt[0] = -a / Aa;
    M.qt[1] = -(a * H.a1_factor) / Aa;
    M.qt[2] = a / Aa;
    M.qt[3] = 2 * 9.81;
    M.qt[4] = g * k1 / Aa;
    M.qt[5] = g * k2 / Aa;
    M.qt[6] = (1 - g) * k2 / Aa;
    M.qt[7] = (1 - g) * k1 / Aa;
  }
  M.t_switch = H.t_switch;
  M.integ_h = H.integ_Ts / (double)H.supersample;
  M.supersample = H.supersample;
  M.nu = H.nu;
}

// ------------------------------------------------------------------------------------------------
// utility kernels (reset!, import/export between the AoS C-ABI layout and the SoA HBM layout)
// ------------------------------------------------------------------------------------------------
struct InitP {
  double mu0[MAX_NX];
  double L0[MAX_NX * MAX_NX];  // row-major lower
};

// reset!(pf)  filtering.jl:4-14: x = xprev ~ initial_density ; (weights handled by Scalars.uniform)
template <int NX>
__global__ void k_init(double* x, long long ld, long long n, long long first, RngKey key, InitP ip) {
  __shared__ MathTab mt;
  math_tab_load(mt);
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double z[NX];
    normals<NX>(key, ST_INIT, 0u, (unsigned long long)(first + i), z, mt);
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c <= r; ++c) acc = fma(ip.L0[r * MAX_NX + c], z[c], acc);
      x[(size_t)r * ld + i] = ip.mu0[r] + acc;
    }
  }
}

__global__ void k_export_x(const double* x, long long ld, long long n, int nx, double* out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    for (int d = 0; d < nx; ++d) out[(size_t)i * nx + d] = x[(size_t)d * ld + i];
}
__global__ void k_import_x(double* x, long long ld, long long n, int nx, const double* in) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    for (int d = 0; d < nx; ++d) x[(size_t)d * ld + i] = in[(size_t)i * nx + d];
}
// weights(pf) / expweights(pf) from the lazy state
__global__ void k_materialise(const double* w, const Scalars* scp, long long n, long long N,
                              double* out_w, double* out_we) {
  const Scalars sc = *scp;
  const double lwN = -log((double)N), lw1N = log(1.0 / (double)N);
  const double inv_s = sc.pend ? 1.0 / sc.pend_s : 1.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double wn, we;
    if (sc.uniform) {
      wn = (sc.uniform == 1) ? lwN : lw1N;
      we = 1.0 / (double)N;
    } else if (sc.pend || sc.stats_ahead) {
      const double wr = w[i];
      if (sc.stats_ahead) {   // APF between predict! and correct!: w is raw, `we` aliases λ in the
        wn = wr;              // reference (filtering.jl:200); we report exp-normalised weights instead
        we = exp(wr - sc.pend_m) * (1.0 / sc.pend_s);
      } else {
        wn = (wr - sc.pend_m) - sc.pend_ls;
        we = exp(wr - sc.pend_m) * inv_s;
      }
    } else {
      wn = w[i];
      we = exp(wn);
    }
    if (out_w) out_w[i] = wn;
    if (out_we) out_we[i] = we;
  }
}
__global__ void k_export_j(const int* j, int identity, long long n, long long first, long long* out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = identity ? (first + i + 1) : ((long long)j[i] + 1);
}
// block partials of (sum we, sum we^2, sum we*x[d]); finished on the host in block order
__global__ void k_wstats(const double* we, const double* x, long long ld, long long n, int nx,
                         double* part /*[grid][2+MAX_NX]*/) {
  __shared__ double sm[NWARP * (2 + MAX_NX)];
  double v[2 + MAX_NX];
  for (int k = 0; k < 2 + MAX_NX; ++k) v[k] = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double e = we[i];
    v[0] += e;
    v[1] = fma(e, e, v[1]);
    for (int d = 0; d < nx; ++d) v[2 + d] = fma(e, x[(size_t)d * ld + i], v[2 + d]);
  }
  for (int k = 0; k < 2 + MAX_NX; ++k)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 2 + MAX_NX; ++k) sm[(threadIdx.x >> 5) * (2 + MAX_NX) + k] = v[k];
  __syncthreads();
  if (threadIdx.x < 2 + MAX_NX) {
    double r = 0.0;
    for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) r += sm[wq * (2 + MAX_NX) + threadIdx.x];
    part[(size_t)blockIdx.x * (2 + MAX_NX) + threadIdx.x] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// stand-alone cooperative kernels at the reference's function boundaries
// ------------------------------------------------------------------------------------------------
// resample(strategy, we, j, bins, M)  resample.jl:17-61 ; we -> bins (scan) -> j (source-side slot ranges)
__global__ void __launch_bounds__(BLOCK)
k_resample(const __grid_constant__ EngineP P, const double* we, double u01, const double* u_slots,
           int M, long long* j_inout, const RangeArg range) {
  __shared__ Shared sh;
  unsigned bar_target = 0;
  long long b = (long long)blockIdx.x * P.chunk, e = b + P.chunk;
  if (e > P.n) e = P.n;
  if (b > P.n) b = P.n;
  double total;
  u64 xs = 0;
  // 1-based ids; slots >= f_total keep the caller's value (resample.jl:26-34)
  // heavy runs are filled by a block partition of the M output slots
  const long long per = ((long long)M + gridDim.x - 1) / gridDim.x;
  const long long s0 = (long long)blockIdx.x * per, s1 = s0 + per;
  (void)resample_indices<long long>(P, sh, (int)b, (int)e, bar_target, [=](int i) { return __ldg(we + i); },
                                    [](int, double v) { return v; }, u01, false, 0u, M, u_slots, j_inout, 1ll, total, xs,
                                    (int)(s0 < M ? s0 : M), (int)(s1 < M ? s1 : M), &range);
}

// resample(ResampleResidual, we, j, bins, M)  resample.jl:63-117 ; the rand() draws of :106 supplied in order
__global__ void __launch_bounds__(BLOCK)
k_resample_residual(const __grid_constant__ EngineP P, const double* we, const double* u_draws, int M,
                    long long* j_inout) {
  __shared__ Shared sh;
  unsigned bar_target = 0;
  long long b = (long long)blockIdx.x * P.chunk, e = b + P.chunk;
  if (e > P.n) e = P.n;
  if (b > P.n) b = P.n;
  const long long per = ((long long)M + gridDim.x - 1) / gridDim.x;
  const long long s0 = (long long)blockIdx.x * per, s1 = s0 + per;
  WeSrc src;
  src.w = we; src.mode = 0;
  src.pm = src.pls = src.inv_s = src.weu = src.wu = 0.0;
  src.T = nullptr; src.hist_w = nullptr; src.hist_we = nullptr;
  double total;
  resample_residual<long long>(P, sh, (int)b, (int)e, bar_target, src, u_draws, 0u, M, j_inout, 1ll, 0,
                               (int)(s0 < M ? s0 : M), (int)(s1 < M ? s1 : M), total);
}

// logsumexp!(w, we)  utils.jl:18-27 on caller-provided arrays
__global__ void __launch_bounds__(BLOCK)
k_logsumexp(const __grid_constant__ EngineP P, double* w, double* we, double* ll_out) {
  __shared__ Shared sh;
  math_tab_load(sh.mt);
  __syncthreads();
  unsigned bar_target = 0;
  long long bl = (long long)blockIdx.x * P.chunk, el = bl + P.chunk;
  if (el > P.n) el = P.n;
  if (bl > P.n) bl = P.n;
  const int beg = (int)bl, end = (int)el;
  Online<1> acc;
  acc.init();
  const double dummy[1] = {0.0};
  for (int i = beg + threadIdx.x; i < end; i += BLOCK) acc.add(w[i], dummy, false, sh.mt);
  u64 xs = 0;
  const Stats st = reduce_stats<1>(P, sh, acc, false, bar_target, xs);
  const double ls = log(st.s), inv = 1.0 / st.s;
  for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
    const double wr = w[i];
    we[i] = exp_nonpos(wr - st.m, sh.mt) * inv;
    w[i] = (wr - st.m) - ls;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *ll_out = st.m + ls;
}

// ------------------------------------------------------------------------------------------------
// the filter handle
// ------------------------------------------------------------------------------------------------
struct llpf_filter;
typedef cudaError_t (*launch_fn)(llpf_filter*, const EngineP&);
typedef cudaError_t (*init_fn)(llpf_filter*, uint64_t epoch);
typedef int (*occ_fn)();

struct llpf_filter {
  llpf_config cfg;
  HostModel hm;
  int device = 0, num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  long long N = 0, n = 0, first = 0, ld = 0;
  // device arena
  char* arena = nullptr;
  size_t arena_bytes = 0;
  double *x[2] = {nullptr, nullptr}, *w = nullptr, *lam = nullptr, *bins = nullptr;
  int* j = nullptr;
  u64* loc = nullptr;
  double* partials = nullptr;
  u64* tots = nullptr;
  unsigned* bar = nullptr;
  Scalars* sc = nullptr;
  double *stage_u = nullptr, *stage_y = nullptr, *wstat = nullptr, *scratch = nullptr;
  double* mbox = nullptr;
  double* bcast = nullptr;
  u64* bcast_flag = nullptr;
  // sharding (one process per GPU): IPC-mapped peer arenas
  int rank = 0, world = 1;
  bool connected = false;
  size_t o_x0 = 0, o_x1 = 0, o_j = 0, o_mbox = 0, o_heavy = 0, o_tots2 = 0, o_pack = 0, o_pack_cnt = 0;
  int pack_stride = 0, pack_state_bytes = 0;
  char* peer_base[MAX_WORLD] = {nullptr};
  // per-run buffers (grow-only)
  double *d_u = nullptr, *d_y = nullptr, *d_ll = nullptr, *d_ess = nullptr, *d_xhat = nullptr;
  int* d_res = nullptr;
  long long cap_T = 0;
  // host mirror
  Scalars hsc;
  Scalars* pin_sc = nullptr;  // pinned staging
  uint64_t epoch = 0;
  long long launches = 0;
  float last_ms = 0.f, last_smooth_ms = 0.f;
  int max_blocks = 1;
  // Ensemble Kalman filter verbs (llpf_enkf_*): state block {ll, t, mean, cov, status}, reduction scratch, inflation
  double* enkf_st = nullptr;
  double* enkf_partials = nullptr;
  double* enkf_out = nullptr;
  size_t enkf_out_doubles = 0;
  int enkf_blocks = 0;
  double enkf_inflation = 1.0;
  launch_fn launch = nullptr;
  init_fn init = nullptr;
  // wide (Float32-particle) engine: device copies of the model in the layouts of WideP
  bool wide = false;
  float *w_At = nullptr, *w_Lt = nullptr, *w_G = nullptr, *w_B = nullptr, *w_mu0 = nullptr, *w_L0 = nullptr;
  double* w_W = nullptr;
  int w_diagL = 0;
  // llpf_run_batch staging (owned by the first handle of a batch; grow-only)
  void* d_batch = nullptr;
  size_t batch_bytes = 0;
  Scalars* d_batch_sc = nullptr;
  int batch_cap = 0;
  // user-defined model (LLPF_DYN_USER): kernel compiled at run time, parameter vector p on the device
  bool user = false;
  cudaKernel_t user_kernel = nullptr;
  double* d_user_p = nullptr;
  int user_np = 0;
};

// the k_engine<NX, NY, DYN, RESID> instantiations live in llpf_engine_inst.cu (one translation unit per group of
// llpf_engine_list.h); this file only sees them as host-side kernel handles
namespace llpf {
const void* engine_kernel_group0(int, int, int, int);
const void* engine_kernel_group1(int, int, int, int);
const void* engine_kernel_group2(int, int, int, int);
const void* engine_kernel_group3(int, int, int, int);
const void* engine_kernel_group4(int, int, int, int);
const void* engine_kernel_group5(int, int, int, int);
const void* engine_kernel_group6(int, int, int, int);
const void* engine_kernel_group7(int, int, int, int);
}  // namespace llpf
static_assert(LLPF_INST_GROUPS == 8, "update the group list");
static const void* engine_kernel(int nx, int ny, int dyn, int resid) {
  typedef const void* (*group_fn)(int, int, int, int);
  static const group_fn groups[] = {engine_kernel_group0, engine_kernel_group1, engine_kernel_group2,
                                    engine_kernel_group3, engine_kernel_group4, engine_kernel_group5,
                                    engine_kernel_group6, engine_kernel_group7};
  for (group_fn g : groups)
    if (const void* k = g(nx, ny, dyn, resid)) return k;
  return nullptr;
}
static int resid_of(const llpf_filter* f);

template <int NX, int NY, int DYN>
static cudaError_t launch_engine(llpf_filter* f, const EngineP& P) {
  ModelP<NX, NY> M;
  fill_modelp<NX, NY>(f->hm, M);
  EngineP Pc = P;
  void* args[] = {(void*)&Pc, (void*)&M};
  const void* k = engine_kernel(NX, NY, DYN, resid_of(f));
  if (!k) return cudaErrorInvalidDeviceFunction;
  return cudaLaunchCooperativeKernel(k, dim3(P.nblocks), dim3(BLOCK), args, 0, f->stream);
}
static int occupancy_engine(int nx, int ny, int dyn, int resid) {
  const void* k = engine_kernel(nx, ny, dyn, resid);
  int occ = 0;
  if (!k || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, BLOCK, 0) != cudaSuccess) return 0;
  return occ;
}
template <int NX>
static cudaError_t launch_init(llpf_filter* f, uint64_t epoch) {
  InitP ip;
  std::memset(&ip, 0, sizeof(ip));
  for (int r = 0; r < NX; ++r) {
    ip.mu0[r] = f->hm.mu0[r];
    for (int c = 0; c <= r; ++c) ip.L0[r * MAX_NX + c] = CMH(f->hm.L0, r, c, NX);
  }
  RngKey key{(uint32_t)f->cfg.seed, (uint32_t)(f->cfg.seed >> 32), (uint32_t)epoch << 8};
  const int grid = (int)std::min<long long>((f->n + 255) / 256, (long long)f->num_sms * 8);
  k_init<NX><<<grid > 0 ? grid : 1, 256, 0, f->stream>>>(f->x[0], f->ld, f->n, f->first, key, ip);
  return cudaGetLastError();
}

// ---- wide engine (llpf_wide.cuh) -----------------------------------------------------------------------
static int upload_wide_model(llpf_filter* f) {
  const HostModel& H = f->hm;
  const int nx = H.nx, ny = H.ny, nu = H.nu;
  std::vector<float> At((size_t)WNX * WNX, 0.f), Lt((size_t)WNX * WNX, 0.f), G((size_t)WNX * WNX, 0.f),
      B((size_t)WNX * MAX_NU, 0.f), mu0(WNX, 0.f), L0((size_t)WNX * WNX, 0.f);
  std::vector<double> W((size_t)ny * ny, 0.0);
  int diag = 1;
  for (int r = 0; r < nx; ++r)
    for (int c = 0; c < nx; ++c) {
      At[(size_t)c * WNX + r] = (float)CMH(H.A, r, c, nx);
      if (c <= r) {
        Lt[(size_t)c * WNX + r] = (float)CMH(H.L1, r, c, nx);
        L0[(size_t)r * WNX + c] = (float)CMH(H.L0, r, c, nx);
        if (c < r && CMH(H.L1, r, c, nx) != 0.0) diag = 0;
      }
    }
  for (int a = 0; a < ny; ++a) {
    for (int c = 0; c < nx; ++c) G[(size_t)a * WNX + c] = (float)CMH(H.G, a, c, ny);
    for (int c = 0; c <= a; ++c) W[(size_t)a * ny + c] = CMH(H.W, a, c, ny);
  }
  for (int r = 0; r < nx; ++r) {
    mu0[r] = (float)H.mu0[r];
    for (int c = 0; c < nu; ++c) B[(size_t)r * MAX_NU + c] = (float)CMH(H.B, r, c, nx);
  }
  f->w_diagL = diag;
  auto up = [&](auto*& dptr, const auto& v) -> cudaError_t {
    typedef typename std::remove_reference<decltype(v)>::type::value_type T;
    if (!dptr) {
      cudaError_t e = cudaMalloc((void**)&dptr, sizeof(T) * v.size());
      if (e != cudaSuccess) return e;
    }
    return cudaMemcpy(dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
  };
  CU(up(f->w_At, At)); CU(up(f->w_Lt, Lt)); CU(up(f->w_G, G)); CU(up(f->w_B, B));
  CU(up(f->w_mu0, mu0)); CU(up(f->w_L0, L0)); CU(up(f->w_W, W));
  return LLPF_OK;
}
static cudaError_t launch_engine_wide(llpf_filter* f, const EngineP& P) {
  WideP Mw;
  std::memset(&Mw, 0, sizeof(Mw));
  Mw.At = f->w_At; Mw.Lt = f->w_Lt; Mw.G = f->w_G; Mw.B = f->w_B; Mw.W = f->w_W;
  Mw.c0 = (float)f->hm.c0;
  Mw.nx = f->hm.nx; Mw.ny = f->hm.ny; Mw.nu = f->hm.nu;
  Mw.diagL = f->w_diagL;
  EngineP Pc = P;
  Pc.want_xhat = 0; Pc.xhat = nullptr;
  void* args[] = {(void*)&Pc, (void*)&Mw};
  return cudaLaunchCooperativeKernel(wide_engine_kernel(), dim3(P.nblocks), dim3(wide_engine_block_threads()), args,
                                     wide_engine_smem_bytes(), f->stream);
}
static int occupancy_engine_wide() {
  if (cudaFuncSetAttribute(wide_engine_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)wide_engine_smem_bytes()) != cudaSuccess)
    return 0;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wide_engine_kernel(), wide_engine_block_threads(),
                                                    wide_engine_smem_bytes()) != cudaSuccess)
    return 0;
  return occ;
}
static cudaError_t launch_init_wide(llpf_filter* f, uint64_t epoch) {
  RngKey key{(uint32_t)f->cfg.seed, (uint32_t)(f->cfg.seed >> 32), (uint32_t)epoch << 8};
  const int grid = (int)std::min<long long>((f->n + 127) / 128, (long long)f->num_sms * 8);
  k_init_wide<<<grid > 0 ? grid : 1, 128, 0, f->stream>>>(reinterpret_cast<float*>(f->x[0]), f->n, f->first, key,
                                                           f->w_mu0, f->w_L0, f->hm.nx);
  return cudaGetLastError();
}

static int resid_of(const llpf_filter* f) {   // which family of engine instantiations: 0 scan-based, 1 residual, 2 Metropolis
  return f->cfg.resampling == LLPF_RESAMPLE_RESIDUAL ? 1 : (f->cfg.resampling == LLPF_RESAMPLE_METROPOLIS ? 2 : 0);
}

// ------------------------------------------------------------------------------------------------
// user-defined models: run-time compilation of k_engine<NX, NY, LLPF_DYN_USER, RESID> with NVRTC
// ------------------------------------------------------------------------------------------------
// libnvrtc is opened lazily (dlopen) so that the library itself has no link-time dependency on it.
namespace {
struct NvrtcApi {
  void* lib = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
  nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
};
NvrtcApi g_nvrtc;
std::mutex g_nvrtc_mu;

bool nvrtc_open(std::string& why) {
  if (g_nvrtc.lib) return true;
  std::vector<std::string> cand;
  if (const char* e = std::getenv("LLPF_NVRTC_PATH")) cand.push_back(e);
  cand.push_back("libnvrtc.so.12");
  cand.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
  cand.push_back("libnvrtc.so");
  void* lib = nullptr;
  for (const std::string& c : cand) {
    lib = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (lib) break;
  }
  if (!lib) { why = "libnvrtc.so.12 not found (set LLPF_NVRTC_PATH)"; return false; }
#define LLPF_SYM(name)                                                        \
  g_nvrtc.name = reinterpret_cast<decltype(g_nvrtc.name)>(dlsym(lib, "nvrtc" #name)); \
  if (!g_nvrtc.name) { why = "libnvrtc lacks nvrtc" #name; dlclose(lib); return false; }
  LLPF_SYM(CreateProgram) LLPF_SYM(DestroyProgram) LLPF_SYM(CompileProgram) LLPF_SYM(GetProgramLogSize)
  LLPF_SYM(GetProgramLog) LLPF_SYM(GetCUBINSize) LLPF_SYM(GetCUBIN) LLPF_SYM(AddNameExpression)
  LLPF_SYM(GetLoweredName) LLPF_SYM(GetErrorString)
#undef LLPF_SYM
  g_nvrtc.lib = lib;
  return true;
}

// directory of this shared library: the engine headers are shipped next to it (csrc/)
std::string library_dir() {
  Dl_info info;
  if (dladdr(reinterpret_cast<const void*>(&llpf_last_error), &info) && info.dli_fname) {
    std::string pth(info.dli_fname);
    const size_t k = pth.find_last_of('/');
    return k == std::string::npos ? std::string(".") : pth.substr(0, k);
  }
  return ".";
}

struct UserKernel { cudaLibrary_t lib; cudaKernel_t kernel; };
std::map<std::string, UserKernel> g_user_kernels;   // key: device ordinal + instantiation + source

// Compiles `#define LLPF_USER_MODEL / #include "llpf_engine.cuh" / <user source> / explicit instantiation` to an sm_100a
// cubin and loads it (cudaLibraryLoadData: context-independent, usable from the handle's device).
int compile_user_kernel(int device, int nx, int ny, int resid, const char* user_src, cudaKernel_t* out) {
  std::lock_guard<std::mutex> lock(g_nvrtc_mu);
  char inst[96];
  std::snprintf(inst, sizeof(inst), "llpf::k_engine<%d, %d, 2, %d>", nx, ny, resid);
  const std::string key = std::to_string(device) + "|" + inst + "|" + user_src;
  auto it = g_user_kernels.find(key);
  if (it != g_user_kernels.end()) { *out = it->second.kernel; return LLPF_OK; }
  std::string why;
  if (!nvrtc_open(why)) return fail(LLPF_ERR_UNSUPPORTED, "user-defined model: " + why);
  std::string src = "#define LLPF_USER_MODEL\n";
  // a source that defines llpf_user::add_noise / correct_state says so by containing this token (llpf.h, llpf_create_user)
  if (std::strstr(user_src, "LLPF_USER_STATE_HOOKS")) src += "#define LLPF_USER_STATE_HOOKS\n";
  src += "#include \"llpf_engine.cuh\"\n#line 1 \"user_model.cu\"\n";
  src += user_src;
  src += "\nnamespace llpf {\ntemplate __global__ void k_engine<" + std::to_string(nx) + ", " + std::to_string(ny) + ", 2, " +
         std::to_string(resid) + ">(const __grid_constant__ EngineP, const __grid_constant__ ModelP<" + std::to_string(nx) +
         ", " + std::to_string(ny) + ">);\n}\n";
  nvrtcProgram prog = nullptr;
  nvrtcResult rc = g_nvrtc.CreateProgram(&prog, src.c_str(), "llpf_user_engine.cu", 0, nullptr, nullptr);
  if (rc != NVRTC_SUCCESS) return fail(LLPF_ERR_CUDA, std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(rc));
  g_nvrtc.AddNameExpression(prog, inst);
  const std::string inc = "-I" + library_dir();
  std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo", inc.c_str(),
                                   "-I/usr/local/cuda/include"};
  rc = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    g_nvrtc.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
    if (log.size() > 6000) log.resize(6000);
    g_nvrtc.DestroyProgram(&prog);
    return fail(LLPF_ERR_BAD_ARG, std::string("user-defined model does not compile (") + g_nvrtc.GetErrorString(rc) + "):\n" + log);
  }
  const char* lowered = nullptr;
  g_nvrtc.GetLoweredName(prog, inst, &lowered);
  size_t nb = 0;
  g_nvrtc.GetCUBINSize(prog, &nb);
  std::vector<char> cubin(nb);
  if (nb) g_nvrtc.GetCUBIN(prog, cubin.data());
  const std::string name = lowered ? lowered : "";
  g_nvrtc.DestroyProgram(&prog);
  if (!nb || name.empty()) return fail(LLPF_ERR_CUDA, "NVRTC produced no cubin / no lowered kernel name");
  UserKernel uk;
  cudaError_t e = cudaLibraryLoadData(&uk.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) return fail(LLPF_ERR_CUDA, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e));
  e = cudaLibraryGetKernel(&uk.kernel, uk.lib, name.c_str());
  if (e != cudaSuccess) return fail(LLPF_ERR_CUDA, std::string("cudaLibraryGetKernel(") + name + "): " + cudaGetErrorString(e));
  g_user_kernels[key] = uk;
  *out = uk.kernel;
  return LLPF_OK;
}
}  // namespace

// the bytes of ModelP<nx, ny> for run-time dimensions (same member order as the template: A, L1, G, W, B, c0, qt[8],
// t_switch, integ_h, supersample, nu)
static std::vector<double> pack_modelp_runtime(const HostModel& H) {
  const int nx = H.nx, ny = H.ny;
  std::vector<double> v((size_t)2 * nx * nx + (size_t)ny * nx + (size_t)ny * ny + (size_t)nx * MAX_NU + 11 + 1, 0.0);
  double* A = v.data();
  double* L1 = A + nx * nx;
  double* G = L1 + nx * nx;
  double* W = G + ny * nx;
  double* B = W + ny * ny;
  double* tail = B + nx * MAX_NU;
  for (int r = 0; r < nx; ++r)
    for (int c = 0; c < nx; ++c) {
      if (!H.A.empty()) A[r * nx + c] = CMH(H.A, r, c, nx);
      L1[r * nx + c] = CMH(H.L1, r, c, nx);
    }
  for (int a = 0; a < ny; ++a) {
    for (int c = 0; c < nx; ++c) G[a * nx + c] = CMH(H.G, a, c, ny);
    for (int c = 0; c < ny; ++c) W[a * ny + c] = CMH(H.W, a, c, ny);
  }
  for (int r = 0; r < nx; ++r)
    for (int c = 0; c < H.nu; ++c)
      if (!H.B.empty()) B[r * MAX_NU + c] = CMH(H.B, r, c, nx);
  tail[0] = H.c0;                               // c0 ; qt[8] stays zero except for the quadtank descriptor
  if (H.dyn == LLPF_DYN_QUADTANK_RK4) {         // as fill_modelp
    const double k1 = H.dynp[1], k2 = H.dynp[2], Aa = H.dynp[3], a = H.dynp[4], g = H.dynp[5];
    const double qt[8] = {-a / Aa, -(a * H.a1_factor) / Aa, a / Aa, 2 * 9.81, g * k1 / Aa, g * k2 / Aa, (1 - g) * k2 / Aa,
                          (1 - g) * k1 / Aa};
    for (int k = 0; k < 8; ++k) tail[1 + k] = qt[k];
  }
  tail[9] = H.t_switch;
  tail[10] = H.integ_Ts / (double)H.supersample;
  int ints[2] = {H.supersample, H.nu};
  std::memcpy(tail + 11, ints, sizeof(ints));
  return v;
}

// EngineP as the LLPF_USER_MODEL build of the engine sees it: one trailing member more
struct EnginePUser {
  EngineP base;
  const double* user_p;
};

static cudaError_t launch_engine_user(llpf_filter* f, const EngineP& P) {
  std::vector<double> M = pack_modelp_runtime(f->hm);
  EnginePUser PU;
  PU.base = P;
  PU.user_p = f->d_user_p;
  void* args[] = {(void*)&PU, (void*)M.data()};
  return cudaLaunchCooperativeKernel((const void*)f->user_kernel, dim3(P.nblocks), dim3(BLOCK), args, 0, f->stream);
}

// host-side half of the dispatch (model packing + init kernel); the device half is llpf_engine_list.h
struct Dispatch {
  int nx, ny, dyn;
  launch_fn launch;
  init_fn init;
};
#define DISP(NX, NY, DYN) \
  { NX, NY, DYN, launch_engine<NX, NY, DYN>, launch_init<NX> }
static const Dispatch g_dispatch[] = {
    DISP(1, 1, 0), DISP(2, 1, 0), DISP(2, 2, 0), DISP(3, 1, 0), DISP(3, 2, 0), DISP(3, 3, 0),
    DISP(4, 1, 0), DISP(4, 2, 0), DISP(4, 3, 0), DISP(4, 4, 0), DISP(6, 2, 0), DISP(6, 3, 0),
    DISP(8, 2, 0), DISP(8, 4, 0), DISP(4, 2, 1),
};

static void push_op(EngineP& P, int kind, int a0, int b0, int count = 1, int da = 0, int db = 0, int flags = 0);
static int launch(llpf_filter* f, const EngineP& P, bool timed);
static const init_fn g_init_by_nx[MAX_NX + 1] = {nullptr, launch_init<1>, launch_init<2>, launch_init<3>, launch_init<4>,
                                                 launch_init<5>, launch_init<6>, launch_init<7>, launch_init<8>};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// staging slot of the step verbs' measurement vectors: y and y1, each up to the widest ny any engine accepts (64: Float32 filters)
constexpr int kStageYDoubles = WNX > 8 ? WNX : 8;

static int check_handle(llpf_handle h) {
  if (!h) return fail(LLPF_ERR_BAD_ARG, "null handle");
  return LLPF_OK;
}

static int sync_scalars(llpf_filter* f) {  // device -> host mirror
  CU(cudaMemcpyAsync(f->pin_sc, f->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  f->hsc = *f->pin_sc;
  return LLPF_OK;
}
static int push_scalars(llpf_filter* f) {  // host mirror -> device
  CU(cudaStreamSynchronize(f->stream));
  *f->pin_sc = f->hsc;
  CU(cudaMemcpyAsync(f->sc, f->pin_sc, sizeof(Scalars), cudaMemcpyHostToDevice, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  return LLPF_OK;
}

extern "C" int llpf_device_count(int* count) {
  if (!count) return fail(LLPF_ERR_BAD_ARG, "null");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(LLPF_ERR_NO_DEVICE, cudaGetErrorString(e));
  }
  *count = c;
  return LLPF_OK;
}

extern "C" int llpf_destroy(llpf_handle h) {
  if (!h) return LLPF_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int r = 0; r < MAX_WORLD; ++r)
    if (r != h->rank && h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
  cudaFree(h->arena);
  cudaFree(h->d_u); cudaFree(h->d_y); cudaFree(h->d_ll); cudaFree(h->d_ess); cudaFree(h->d_xhat);
  cudaFree(h->d_res);
  cudaFree(h->w_At); cudaFree(h->w_Lt); cudaFree(h->w_G); cudaFree(h->w_B); cudaFree(h->w_mu0); cudaFree(h->w_L0);
  cudaFree(h->w_W);
  cudaFree(h->d_user_p);
  cudaFree(h->d_batch); cudaFree(h->d_batch_sc);
  cudaFree(h->enkf_st); cudaFree(h->enkf_partials); cudaFree(h->enkf_out);
  if (h->pin_sc) cudaFreeHost(h->pin_sc);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return LLPF_OK;
}

extern "C" int llpf_set_model(llpf_handle h, const llpf_model* model) {
  OKR(check_handle(h));
  HostModel hm;
  OKR(build_host_model(model, hm, h->wide));
  if (hm.nx != h->hm.nx || hm.ny != h->hm.ny || hm.nu != h->hm.nu || hm.dyn != h->hm.dyn)
    return fail(LLPF_ERR_BAD_ARG, "llpf_set_model cannot change dimensions or dynamics kind");
  h->hm = hm;
  if (h->wide) {
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    OKR(upload_wide_model(h));
  }
  return LLPF_OK;
}

extern "C" int llpf_reset(llpf_handle h, uint64_t epoch);

static int create_impl(const llpf_config* cfg, const llpf_model* model, const char* user_src, const double* user_p,
                       int32_t user_np, llpf_handle* out) {
  if (!cfg || !model || !out) return fail(LLPF_ERR_BAD_ARG, "null argument");
  *out = nullptr;
  const bool user = user_src != nullptr;
  if (user != (model->dynamics == LLPF_DYN_USER))
    return fail(LLPF_ERR_BAD_ARG, "LLPF_DYN_USER models are created with llpf_create_user (and only those)");
  if (user && (user_np < 0 || (user_np > 0 && !user_p))) return fail(LLPF_ERR_BAD_ARG, "bad parameter vector");
  if (cfg->N < 1 || cfg->N >= (1ll << 31)) return fail(LLPF_ERR_BAD_ARG, "need 1 <= N < 2^31");
  if (cfg->filter < 0 || cfg->filter > 3) return fail(LLPF_ERR_BAD_ARG, "unknown filter kind");
  if (user && std::strstr(user_src, "LLPF_USER_STATE_HOOKS") && cfg->filter >= LLPF_FILTER_AUX)
    return fail(LLPF_ERR_UNSUPPORTED, "user models with state hooks (add_noise / correct_state, e.g. RBPF) run as ParticleFilter / "
                                      "AdvancedParticleFilter: the auxiliary filter's two-stage step does not call the hooks");
  if (cfg->resampling != LLPF_RESAMPLE_SYSTEMATIC && cfg->resampling != LLPF_RESAMPLE_STRATIFIED &&
      cfg->resampling != LLPF_RESAMPLE_RESIDUAL && cfg->resampling != LLPF_RESAMPLE_METROPOLIS)
    return fail(LLPF_ERR_BAD_ARG, "unknown resampling strategy");
  const int world = cfg->world < 1 ? 1 : cfg->world;
  if (world > 1 && (cfg->resampling == LLPF_RESAMPLE_RESIDUAL || cfg->resampling == LLPF_RESAMPLE_METROPOLIS))
    return fail(LLPF_ERR_UNSUPPORTED, "residual / Metropolis resampling are single-GPU (sharded filters: systematic or stratified)");
  if (cfg->resampling == LLPF_RESAMPLE_METROPOLIS && cfg->particle_dtype == LLPF_PARTICLE_F32)
    return fail(LLPF_ERR_UNSUPPORTED, "Metropolis resampling: Float64-particle filters");
  if (cfg->metropolis_steps < 0) return fail(LLPF_ERR_BAD_ARG, "metropolis_steps must be >= 0");
  if (world > MAX_WORLD) return fail(LLPF_ERR_UNSUPPORTED, "at most 8 ranks (one box)");
  if (world > 1) {
    if (cfg->rank < 0 || cfg->rank >= world) return fail(LLPF_ERR_BAD_ARG, "bad rank");
    if (cfg->N % world) return fail(LLPF_ERR_BAD_ARG, "N must be divisible by the number of ranks");
    if (cfg->scan_mode != LLPF_SCAN_FAST) return fail(LLPF_ERR_UNSUPPORTED, "sharded filters use the fixed-point scan (serial cumsum is single-GPU)");
  }
  int ndev = 0;
  OKR(llpf_device_count(&ndev));
  if (ndev < 1) return fail(LLPF_ERR_NO_DEVICE, "no CUDA device; the product path has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(LLPF_ERR_BAD_ARG, "bad device ordinal");

  if (cfg->particle_dtype != LLPF_PARTICLE_F64 && cfg->particle_dtype != LLPF_PARTICLE_F32)
    return fail(LLPF_ERR_BAD_ARG, "unknown particle_dtype");
  const bool wide = cfg->particle_dtype == LLPF_PARTICLE_F32;
  if (wide && cfg->filter != LLPF_FILTER_PF && cfg->filter != LLPF_FILTER_ADVANCED)
    return fail(LLPF_ERR_UNSUPPORTED, "Float32 particles: ParticleFilter / AdvancedParticleFilter only");
  llpf_filter* f = new llpf_filter();
  f->cfg = *cfg;
  f->cfg.world = world;
  f->wide = wide;
  int rc = build_host_model(model, f->hm, wide);
  if (rc) { delete f; return rc; }
  const Dispatch* d = nullptr;
  f->user = user;
  if (user) {
    if (wide) { delete f; return fail(LLPF_ERR_UNSUPPORTED, "user-defined models: Float64 particles only"); }
    if (cudaSetDevice(cfg->device) != cudaSuccess) { delete f; return fail(LLPF_ERR_CUDA, "cudaSetDevice"); }
    rc = compile_user_kernel(cfg->device, f->hm.nx, f->hm.ny, resid_of(f), user_src, &f->user_kernel);
    if (rc) { delete f; return rc; }
    f->launch = launch_engine_user;
    f->init = g_init_by_nx[f->hm.nx];
  } else if (!wide) {
    for (const Dispatch& e : g_dispatch)
      if (e.nx == f->hm.nx && e.ny == f->hm.ny && e.dyn == f->hm.dyn) d = &e;
    if (d && !engine_kernel(d->nx, d->ny, d->dyn, resid_of(f))) d = nullptr;
    if (!d) {
      delete f;
      return fail(LLPF_ERR_UNSUPPORTED, "no kernel instantiated for this (nx, ny, dynamics)");
    }
    f->launch = d->launch;
    f->init = d->init;
  } else {
    f->launch = launch_engine_wide;
    f->init = launch_init_wide;
  }
  f->device = cfg->device;
#define CUF(expr)                                                                       \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      llpf_destroy(f);                                                                  \
      return fail(LLPF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));   \
    }                                                                                   \
  } while (0)
  CUF(cudaSetDevice(f->device));
  cudaDeviceProp prop;
  CUF(cudaGetDeviceProperties(&prop, f->device));
  f->num_sms = prop.multiProcessorCount;
  if (!prop.cooperativeLaunch) {
    llpf_destroy(f);
    return fail(LLPF_ERR_UNSUPPORTED, "device lacks cooperative launch");
  }
  int occ = 0;
  if (user) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)f->user_kernel, BLOCK, 0) != cudaSuccess) occ = 0;
  } else {
    occ = wide ? occupancy_engine_wide() : occupancy_engine(d->nx, d->ny, d->dyn, resid_of(f));
  }
  if (occ < 1) {
    llpf_destroy(f);
    return fail(LLPF_ERR_CUDA, "engine kernel does not fit on an SM (is this an sm_100a device?)");
  }
  f->max_blocks = std::min(occ * f->num_sms, MAX_BLOCKS - 1);
  CUF(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  CUF(cudaEventCreate(&f->ev0));
  CUF(cudaEventCreate(&f->ev1));
  CUF(cudaMallocHost(&f->pin_sc, sizeof(Scalars)));

  f->N = cfg->N;
  f->world = world;
  f->rank = world > 1 ? cfg->rank : 0;
  f->n = cfg->N / world;
  f->first = (long long)f->rank * f->n;
  f->ld = (long long)align_up((size_t)f->n, 32);
  const int nx = f->hm.nx;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  // particles: SoA f64 [nx][ld], or (wide) AoS f32 [n][64]
  const size_t xbytes = wide ? (size_t)WNX * f->ld * 4 : (size_t)nx * f->ld * 8;
  const size_t o_x0 = take(xbytes), o_x1 = take(xbytes);
  const size_t o_w = take((size_t)f->ld * 8), o_lam = take((size_t)f->ld * 8), o_bins = take((size_t)f->ld * 8);
  const size_t o_j = take((size_t)f->ld * 4);
  const size_t o_loc = take((size_t)f->ld * 8);
  const size_t o_part = take((size_t)MAX_BLOCKS * PS * 8), o_tots = take((size_t)MAX_BLOCKS * 8);
  const size_t o_bar = take((size_t)BAR_TOTAL_WORDS * 4), o_sc = take(sizeof(Scalars));
  const size_t o_su = take(2 * MAX_NU * 8), o_sy = take(2 * kStageYDoubles * 8), o_ws = take(256 * (2 + WNX) * 8);
  const size_t o_scr = take((size_t)(nx + 2) * f->ld * 8);
  const size_t o_mbox = take((size_t)2 * MAX_WORLD * MBOX_WORDS * 8);
  const size_t o_bcast = take((size_t)2 * MAX_WORLD * MBOX_DOUBLES * 8), o_bflag = take(64);
  const size_t o_heavy = take((size_t)(1 + 3 * HEAVY_MAX) * 4);
  f->o_tots2 = take((size_t)MAX_BLOCKS * 8);
  f->o_pack_cnt = take((size_t)MAX_WORLD * 4);
  if (world > 1) {
    // packed particle exchange of sharded resampling: one region per source rank, one entry per local particle at most
    // (a particle contributes at most one entry per destination); entry = state padded to 16 B + int4 meta
    f->pack_state_bytes = wide ? WNX * 4 : (int)align_up((size_t)nx * 8, 16);
    f->pack_stride = wide ? WNX * 4 + 32 : f->pack_state_bytes + 16;   // wide rows stay 32-byte aligned (ld.v8.f32)
    f->o_pack = take((size_t)world * (size_t)f->n * (size_t)f->pack_stride);
  }
  f->o_x0 = o_x0; f->o_x1 = o_x1; f->o_j = o_j; f->o_mbox = o_mbox; f->o_heavy = o_heavy;
  f->arena_bytes = off;
  CUF(cudaMalloc(&f->arena, f->arena_bytes));
  CUF(cudaMemsetAsync(f->arena, 0, f->arena_bytes, f->stream));
  f->x[0] = (double*)(f->arena + o_x0); f->x[1] = (double*)(f->arena + o_x1);
  f->w = (double*)(f->arena + o_w); f->lam = (double*)(f->arena + o_lam);
  f->bins = (double*)(f->arena + o_bins); f->j = (int*)(f->arena + o_j);
  f->loc = (u64*)(f->arena + o_loc);
  f->partials = (double*)(f->arena + o_part); f->tots = (u64*)(f->arena + o_tots);
  f->bar = (unsigned*)(f->arena + o_bar); f->sc = (Scalars*)(f->arena + o_sc);
  f->stage_u = (double*)(f->arena + o_su); f->stage_y = (double*)(f->arena + o_sy);
  f->wstat = (double*)(f->arena + o_ws); f->scratch = (double*)(f->arena + o_scr);
  f->mbox = (double*)(f->arena + o_mbox);
  f->bcast = (double*)(f->arena + o_bcast); f->bcast_flag = (u64*)(f->arena + o_bflag);
  f->peer_base[f->rank] = f->arena;
#undef CUF
  if (wide) {
    rc = upload_wide_model(f);
    if (rc) { llpf_destroy(f); return rc; }
  }
  if (user) {
    f->user_np = user_np;
    if (cudaMalloc(&f->d_user_p, sizeof(double) * (user_np > 0 ? user_np : 1)) != cudaSuccess ||
        (user_np > 0 && cudaMemcpy(f->d_user_p, user_p, sizeof(double) * user_np, cudaMemcpyHostToDevice) != cudaSuccess)) {
      llpf_destroy(f);
      return fail(LLPF_ERR_CUDA, "allocating the user parameter vector");
    }
  }
  rc = llpf_reset(f, 0);
  if (rc) { llpf_destroy(f); return rc; }
  f->hsc.t_index = 0;  // PFstate(...,Ref(0)) PFtypes.jl:70 ; reset! sets 1
  rc = push_scalars(f);
  if (rc) { llpf_destroy(f); return rc; }
  *out = f;
  return LLPF_OK;
}

// the scalar state right after reset!  (filtering.jl:4-14)
static Scalars reset_scalars(const llpf_filter* h) {
  Scalars s;
  std::memset(&s, 0, sizeof(s));
  s.t_index = 1;          // filtering.jl:13
  s.cur = 0;
  s.uniform = 1;          // w = -log N, we = 1/N   filtering.jl:11-12
  s.stats_valid = 1;
  s.ess = (double)h->N;
  s.j_identity = 1;
  s.xseq = h->hsc.xseq;   // the peer-exchange counter runs on across resets (mailboxes are never re-zeroed)
  return s;
}

extern "C" int llpf_create(const llpf_config* cfg, const llpf_model* model, llpf_handle* out) {
  return create_impl(cfg, model, nullptr, nullptr, 0, out);
}
extern "C" int llpf_create_user(const llpf_config* cfg, const llpf_model* model, const char* cuda_source,
                                const double* p, int32_t np, llpf_handle* out) {
  if (!cuda_source) return fail(LLPF_ERR_BAD_ARG, "cuda_source is null");
  return create_impl(cfg, model, cuda_source, p, np, out);
}
extern "C" int llpf_set_user_params(llpf_handle h, const double* p, int32_t np) {
  OKR(check_handle(h));
  if (!h->user) return fail(LLPF_ERR_BAD_ARG, "not a user-defined model");
  if (np != h->user_np || (np > 0 && !p)) return fail(LLPF_ERR_BAD_ARG, "parameter vector length differs from the one given at creation");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  if (np > 0) CU(cudaMemcpy(h->d_user_p, p, sizeof(double) * np, cudaMemcpyHostToDevice));
  return LLPF_OK;
}

extern "C" int llpf_reset(llpf_handle h, uint64_t epoch) {
  OKR(check_handle(h));
  CU(cudaSetDevice(h->device));
  h->epoch = epoch;
  CU(h->init(h, epoch));
  h->launches += 1;
  h->hsc = reset_scalars(h);
  return push_scalars(h);
}

// ------------------------------------------------------------------------------------------------
// launching the engine
// ------------------------------------------------------------------------------------------------
static void base_params(llpf_filter* f, EngineP& P) {
  std::memset(&P, 0, sizeof(P));
  P.x[0] = f->x[0]; P.x[1] = f->x[1]; P.ld = f->ld;
  P.w = f->w; P.lam = f->lam; P.bins = f->bins; P.j = f->j; P.loc = f->loc;
  P.bar = f->bar; P.partials = f->partials; P.tots = f->tots; P.sc = f->sc;
  P.N = f->N; P.n = (int)f->n; P.first = (int)f->first;
  P.filter = f->cfg.filter;
  P.Ts = f->cfg.Ts;
  P.thr = f->cfg.resample_threshold;
  P.strategy = f->cfg.resampling;
  P.scan_mode = f->cfg.scan_mode;
  long long nb = (f->n + BLOCK - 1) / BLOCK;
  if (nb > f->max_blocks) nb = f->max_blocks;
  if (nb < 1 || f->cfg.single_block) nb = 1;   // single_block: PMMH-sized filters, batchable (llpf_run_batch)
  P.nblocks = (int)nb;
  P.chunk = (int)((f->n + nb - 1) / nb);
  P.chunk = (P.chunk + 1) & ~1;   // even chunk starts: 16-byte aligned particle pairs in the scan
  P.key = RngKey{(uint32_t)f->cfg.seed, (uint32_t)(f->cfg.seed >> 32), (uint32_t)f->epoch << 8};
  P.rank = f->rank; P.world = f->world;
  P.fix_scale = FIX_SCALE; P.fix_inv = FIX_INV;
  P.bcast = f->bcast; P.bcast_flag = f->bcast_flag;
  P.heavy = (int*)(f->arena + f->o_heavy);
  P.tots2 = (u64*)(f->arena + f->o_tots2);
  P.nx = f->hm.nx; P.wide = f->wide ? 1 : 0;
  P.metro_steps = f->cfg.metropolis_steps;
  P.pack_cnt = (int*)(f->arena + f->o_pack_cnt);
  P.pack_cap = f->n; P.pack_stride = f->pack_stride; P.pack_state_bytes = f->pack_state_bytes;
  P.pack_in = f->world > 1 ? f->arena + f->o_pack : nullptr;
  for (int r = 0; r < f->world; ++r) {
    char* base = f->peer_base[r];
    P.peer_x[r][0] = base ? (double*)(base + f->o_x0) : nullptr;
    P.peer_x[r][1] = base ? (double*)(base + f->o_x1) : nullptr;
    P.peer_j[r] = base ? (int*)(base + f->o_j) : nullptr;
    P.peer_mbox[r] = base ? (double*)(base + f->o_mbox) : nullptr;
    P.peer_heavy[r] = base ? (int*)(base + f->o_heavy) : nullptr;
    P.peer_pack[r] = (base && f->world > 1) ? base + f->o_pack : nullptr;
  }
}

// ---- op-list construction (the kernel executes EngineP::ops in order) -------------------------------
static void push_op(EngineP& P, int kind, int a0, int b0, int count, int da, int db, int flags) {
  if (count < 1 || P.nops >= MAX_OPS) return;
  P.ops[P.nops++] = OpRun{kind, a0, b0, count, da, db, flags, 0};
}
// correct!(pfa) = logsumexp! only (filtering.jl:170-174): reuse the stats aux_step left behind when they
// describe the current weights, otherwise a reduce-only sweep
static void push_aux_correct(llpf_filter* f, EngineP& P, int k) {
  if (f->hsc.stats_ahead) push_op(P, OP_AUX_CSTATS, k, 0);
  else push_op(P, OP_PF, 0, k, 1, 0, 0, OPF_SKIP_MEAS);
}

static int launch(llpf_filter* f, const EngineP& P, bool timed) {
  if (f->world > 1 && !f->connected)
    return fail(LLPF_ERR_BAD_ARG, "sharded filter: call llpf_shard_connect with every rank's blob first");
  CU(cudaMemsetAsync(f->bar, 0, sizeof(unsigned) * BAR_TOTAL_WORDS, f->stream));
  if (timed) CU(cudaEventRecord(f->ev0, f->stream));
  CU(f->launch(f, P));
  f->launches += 1;
  if (timed) CU(cudaEventRecord(f->ev1, f->stream));
  OKR(sync_scalars(f));
  if (timed) CU(cudaEventElapsedTime(&f->last_ms, f->ev0, f->ev1));
  return LLPF_OK;
}

static int stage_inputs(llpf_filter* f, const double* u, const double* y0, const double* y1) {
  const int nu = f->hm.nu, ny = f->hm.ny;
  if (ny > kStageYDoubles || nu > MAX_NU) return fail(LLPF_ERR_BAD_ARG, "u / y do not fit the staging slots");
  if (nu > 0) {
    if (!u) return fail(LLPF_ERR_BAD_ARG, "u is null");
    CU(cudaMemcpyAsync(f->stage_u, u, sizeof(double) * nu, cudaMemcpyHostToDevice, f->stream));
  }
  if (y0) CU(cudaMemcpyAsync(f->stage_y, y0, sizeof(double) * ny, cudaMemcpyHostToDevice, f->stream));
  if (y1) CU(cudaMemcpyAsync(f->stage_y + ny, y1, sizeof(double) * ny, cudaMemcpyHostToDevice, f->stream));
  return LLPF_OK;
}

static bool is_aux(const llpf_filter* f) { return f->cfg.filter >= LLPF_FILTER_AUX; }

static int finish_ll(llpf_filter* f, double* ll) {
  if (ll) *ll = f->hsc.ll_last;
  if (f->hsc.nonfinite) {
    f->hsc.nonfinite = 0;
    OKR(push_scalars(f));
    return fail(LLPF_ERR_NONFINITE, "log-likelihood is not finite (weight collapse)");
  }
  return LLPF_OK;
}

extern "C" int llpf_correct(llpf_handle h, const double* u, const double* y, double t, double* ll) {
  OKR(check_handle(h));
  if (!y) return fail(LLPF_ERR_BAD_ARG, "y is null");
  CU(cudaSetDevice(h->device));
  OKR(stage_inputs(h, u, y, nullptr));
  EngineP P;
  base_params(h, P);
  P.u = h->stage_u; P.y = h->stage_y;
  P.use_t_override = 1; P.t_override = t; P.want_xhat = 1;
  if (is_aux(h)) push_aux_correct(h, P, 1);
  else push_op(P, OP_PF, 0, 1);
  OKR(launch(h, P, false));
  return finish_ll(h, ll);
}

extern "C" int llpf_predict(llpf_handle h, const double* u, double t) {
  OKR(check_handle(h));
  if (is_aux(h)) return fail(LLPF_ERR_BAD_ARG, "AuxiliaryParticleFilter: use llpf_predict_aux(u, y1, t)");
  CU(cudaSetDevice(h->device));
  OKR(stage_inputs(h, u, nullptr, nullptr));
  EngineP P;
  base_params(h, P);
  P.u = h->stage_u; P.y = h->stage_y;
  P.use_t_override = 1; P.t_override = t; P.want_xhat = 0;
  if (!h->hsc.stats_valid) push_op(P, OP_PF, 0, 1, 1, 0, 0, OPF_SKIP_MEAS);  // refresh ESS (resample.jl:5-10)
  push_op(P, OP_PF, 1, 0);
  return launch(h, P, false);
}

extern "C" int llpf_predict_aux(llpf_handle h, const double* u, const double* y1, double t) {
  OKR(check_handle(h));
  if (!is_aux(h)) return fail(LLPF_ERR_BAD_ARG, "llpf_predict_aux needs an AuxiliaryParticleFilter");
  if (!y1) return fail(LLPF_ERR_BAD_ARG, "y1 is null");
  CU(cudaSetDevice(h->device));
  OKR(stage_inputs(h, u, y1, nullptr));
  EngineP P;
  base_params(h, P);
  P.u = h->stage_u; P.y = h->stage_y;
  P.use_t_override = 1; P.t_override = t; P.want_xhat = 1;
  push_op(P, OP_AUX_STEP, 1, 1);   // y1 staged at y[0]
  return launch(h, P, false);
}

extern "C" int llpf_update(llpf_handle h, const double* u, const double* y, const double* y1, double t,
                           double* ll) {
  OKR(check_handle(h));
  if (!y) return fail(LLPF_ERR_BAD_ARG, "y is null");
  CU(cudaSetDevice(h->device));
  EngineP P;
  base_params(h, P);
  P.u = h->stage_u; P.y = h->stage_y;
  P.use_t_override = 1; P.t_override = t; P.want_xhat = 1;
  if (is_aux(h)) {
    if (!y1) return fail(LLPF_ERR_BAD_ARG, "update!(pfa,u,y,y1): y1 is null");
    OKR(stage_inputs(h, u, y, y1));
    push_aux_correct(h, P, 1);
    push_op(P, OP_AUX_STEP, 1, 2);   // y1 staged at y[1]
  } else {
    OKR(stage_inputs(h, u, y, nullptr));
    push_op(P, OP_PF, 0, 1);
    push_op(P, OP_PF, 1, 0);
  }
  OKR(launch(h, P, false));
  return finish_ll(h, ll);
}

// ------------------------------------------------------------------------------------------------
// trajectory drivers
// ------------------------------------------------------------------------------------------------
static int ensure_run_buffers(llpf_filter* f, long long T) {
  if (T <= f->cap_T) return LLPF_OK;
  cudaFree(f->d_u); cudaFree(f->d_y); cudaFree(f->d_ll); cudaFree(f->d_ess); cudaFree(f->d_xhat);
  cudaFree(f->d_res);
  f->d_u = f->d_y = f->d_ll = f->d_ess = f->d_xhat = nullptr; f->d_res = nullptr; f->cap_T = 0;
  const int nu = f->hm.nu > 0 ? f->hm.nu : 1;
  CU(cudaMalloc(&f->d_u, sizeof(double) * T * nu));
  CU(cudaMalloc(&f->d_y, sizeof(double) * T * f->hm.ny));
  CU(cudaMalloc(&f->d_ll, sizeof(double) * T));
  CU(cudaMalloc(&f->d_ess, sizeof(double) * T));
  CU(cudaMalloc(&f->d_xhat, sizeof(double) * T * f->hm.nx));
  CU(cudaMalloc(&f->d_res, sizeof(int) * T));
  f->cap_T = T;
  return LLPF_OK;
}

struct DevHistory {   // forward history left on the device for a consumer (the smoother); owner frees
  double *x = nullptr, *w = nullptr, *we = nullptr;
  void release() { cudaFree(x); cudaFree(w); cudaFree(we); x = w = we = nullptr; }
};
struct HistoryGuard {   // frees the history buffers of a run on every exit path unless they were handed over
  double *&x, *&w, *&we;
  ~HistoryGuard() { cudaFree(x); cudaFree(w); cudaFree(we); }
};

// the op list of a whole trajectory: forward_trajectory (filtering.jl:343-384) or loglik (smoothing.jl:227-236)
static void build_run_ops(llpf_filter* f, EngineP& P, int Ti, int32_t time_convention) {
  // APF passes t=(k-1)*Ts explicitly in both drivers (filtering.jl:376, smoothing.jl:234-235)
  P.time_conv = is_aux(f) ? 0 : (time_convention == LLPF_TIME_LOGLIK ? 1 : 0);
  if (!is_aux(f)) {
    // W(1) [P(k)+W(k+1)] k=1..T-1  P(T)
    push_op(P, OP_PF, 0, 1);
    push_op(P, OP_PF, 1, 2, Ti - 1, 1, 1);
    push_op(P, OP_PF, Ti, 0);
  } else if (time_convention != LLPF_TIME_LOGLIK) {
    // forward_trajectory(pfa) filtering.jl:367-384: correct!(k); k<T && predict!(k, y[k+1])
    push_aux_correct(f, P, 1);
    push_op(P, OP_AUX_STEP, 1, 2, Ti - 1, 1, 1, OPF_POST_CSTATS);
    push_op(P, OP_FLUSH_WHIST, Ti, 0);
  } else {
    // loglik(pfa) smoothing.jl:232-236: T-1 update!(pfa,u,y,y1) then the INNER filter's update! on the last
    // sample: its correct! adds the likelihood of y[T] on top of the raw w = lam - log N, then predict!
    if (Ti > 1) {
      push_aux_correct(f, P, 1);
      push_op(P, OP_AUX_STEP, 1, 2, Ti - 2, 1, 1, OPF_POST_CSTATS);
      push_op(P, OP_AUX_STEP, Ti - 1, Ti);
      push_op(P, OP_PF, 0, Ti, 1, 0, 0, OPF_RAW_WEIGHTS);
    } else {
      push_op(P, OP_PF, 0, 1);
    }
    push_op(P, OP_PF, Ti, 0);
  }
}

static int run_impl(llpf_filter* f, long long T, const double* u_dev, const double* y_dev,
                    int32_t time_convention, uint64_t epoch, double* ll, const llpf_run_outputs* out,
                    DevHistory* keep = nullptr) {
  if (T < 1 || T > (1ll << 30)) return fail(LLPF_ERR_BAD_ARG, "bad T");
  if (f->wide && out && (out->xhat || out->x_hist || out->w_hist || out->we_hist))
    return fail(LLPF_ERR_UNSUPPORTED, "Float32-particle filters: per-step xhat and the x/w/we history are not recorded "
                                      "in the fused loop (use the step verbs + accessors)");
  OKR(llpf_reset(f, epoch));
  EngineP P;
  base_params(f, P);
  P.u = u_dev; P.y = y_dev;
  build_run_ops(f, P, (int)T, time_convention);
  double *xh = nullptr, *wh = nullptr, *weh = nullptr;
  HistoryGuard hist_guard{xh, wh, weh};
  const size_t NT = (size_t)f->N * (size_t)T;
  if (out) {
    P.ll_steps = out->ll_steps ? f->d_ll : nullptr;
    P.ess_steps = out->ess_steps ? f->d_ess : nullptr;
    P.resampled = out->resampled ? f->d_res : nullptr;
    P.xhat = out->xhat ? f->d_xhat : nullptr;
    P.want_xhat = out->xhat ? 1 : 0;
    if (out->x_hist) { CU(cudaMalloc(&xh, NT * f->hm.nx * 8)); P.x_hist = xh; }
    if (out->w_hist || out->we_hist) {
      CU(cudaMalloc(&wh, NT * 8));
      CU(cudaMalloc(&weh, NT * 8));
      P.w_hist = wh; P.we_hist = weh;
    }
    if (P.resampled) CU(cudaMemsetAsync(f->d_res, 0, sizeof(int) * T, f->stream));
  }
  if (keep) {
    if (!xh) { CU(cudaMalloc(&xh, NT * f->hm.nx * 8)); P.x_hist = xh; }
    if (!wh) {
      CU(cudaMalloc(&wh, NT * 8));
      CU(cudaMalloc(&weh, NT * 8));
      P.w_hist = wh; P.we_hist = weh;
    }
  }
#ifdef LLPF_PHASE_TIMING
  long long* d_dbg = nullptr;
  const char* dump = std::getenv("LLPF_PHASE_DUMP");
  if (dump) {
    const size_t dbg_words = (size_t)(16 + 4 * MAX_BLOCKS) * (T + 2);   // block 0's phases + per-block (pass start, sweep end, stats seen, indices done)
    CU(cudaMalloc(&d_dbg, sizeof(long long) * dbg_words));
    CU(cudaMemsetAsync(d_dbg, 0, sizeof(long long) * dbg_words, f->stream));
    P.dbg = d_dbg;
    P.dbg_T = (int)T;
  }
#endif
  int rc = launch(f, P, true);
#ifdef LLPF_PHASE_TIMING
  if (dump && rc == LLPF_OK) {
    std::vector<long long> hb((size_t)(16 + 4 * MAX_BLOCKS) * (T + 2));
    cudaMemcpy(hb.data(), d_dbg, sizeof(long long) * hb.size(), cudaMemcpyDeviceToHost);
    FILE* fp = std::fopen(dump, "wb");
    if (fp) { std::fwrite(hb.data(), sizeof(long long), hb.size(), fp); std::fclose(fp); }
  }
  cudaFree(d_dbg);
#endif
  if (rc == LLPF_OK && out) {
    cudaError_t e = cudaSuccess;
    if (out->ll_steps) e = cudaMemcpyAsync(out->ll_steps, f->d_ll, 8 * T, cudaMemcpyDeviceToHost, f->stream);
    if (!e && out->ess_steps) e = cudaMemcpyAsync(out->ess_steps, f->d_ess, 8 * T, cudaMemcpyDeviceToHost, f->stream);
    if (!e && out->resampled) e = cudaMemcpyAsync(out->resampled, f->d_res, 4 * T, cudaMemcpyDeviceToHost, f->stream);
    if (!e && out->xhat) e = cudaMemcpyAsync(out->xhat, f->d_xhat, 8 * T * f->hm.nx, cudaMemcpyDeviceToHost, f->stream);
    if (!e && out->x_hist) e = cudaMemcpyAsync(out->x_hist, xh, NT * f->hm.nx * 8, cudaMemcpyDeviceToHost, f->stream);
    if (!e && out->w_hist) e = cudaMemcpyAsync(out->w_hist, wh, NT * 8, cudaMemcpyDeviceToHost, f->stream);
    if (!e && out->we_hist) e = cudaMemcpyAsync(out->we_hist, weh, NT * 8, cudaMemcpyDeviceToHost, f->stream);
    if (!e) e = cudaStreamSynchronize(f->stream);
    if (e) rc = fail(LLPF_ERR_CUDA, std::string("copying run outputs: ") + cudaGetErrorString(e));
  }
  if (keep && rc == LLPF_OK) {
    keep->x = xh; keep->w = wh; keep->we = weh;
    xh = wh = weh = nullptr;   // ownership moved to the caller
  }
  if (rc) return rc;
  if (ll) *ll = f->hsc.ll_total;
  if (f->hsc.nonfinite) return fail(LLPF_ERR_NONFINITE, "log-likelihood is not finite (weight collapse)");
  return LLPF_OK;
}

extern "C" int llpf_run(llpf_handle h, int64_t T, const double* u, const double* y,
                        int32_t time_convention, uint64_t epoch, double* ll, const llpf_run_outputs* out) {
  OKR(check_handle(h));
  if (!y || (h->hm.nu > 0 && !u)) return fail(LLPF_ERR_BAD_ARG, "u / y is null");
  if (T < 1) return fail(LLPF_ERR_BAD_ARG, "bad T");
  CU(cudaSetDevice(h->device));
  OKR(ensure_run_buffers(h, T));
  if (h->hm.nu > 0)
    CU(cudaMemcpyAsync(h->d_u, u, sizeof(double) * T * h->hm.nu, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_y, y, sizeof(double) * T * h->hm.ny, cudaMemcpyHostToDevice, h->stream));
  return run_impl(h, T, h->d_u, h->d_y, time_convention, epoch, ll, out);
}

extern "C" int llpf_run_dev(llpf_handle h, int64_t T, const double* u_dev, const double* y_dev,
                            int32_t time_convention, uint64_t epoch, double* ll, const llpf_run_outputs* out) {
  OKR(check_handle(h));
  if (!y_dev || (h->hm.nu > 0 && !u_dev)) return fail(LLPF_ERR_BAD_ARG, "u_dev / y_dev is null");
  if (T < 1) return fail(LLPF_ERR_BAD_ARG, "bad T");
  CU(cudaSetDevice(h->device));
  OKR(ensure_run_buffers(h, T));
  return run_impl(h, T, u_dev, y_dev, time_convention, epoch, ll, out);
}


// ------------------------------------------------------------------------------------------------
// batched multi-chain loglik: one launch, one thread block per filter (llpf_engine_batch.cu)
// ------------------------------------------------------------------------------------------------
namespace llpf {
const void* engine_batch_kernel(int nx, int ny, int dyn, int resid);
}

template <int NX, int NY>
static int run_batch_t(int C, llpf_filter* const* fs, long long T, int32_t conv, const uint64_t* epochs, double* ll_out,
                       const void* kernel) {
  typedef BatchItem<NX, NY> Item;
  llpf_filter* f0 = fs[0];
  const size_t bytes = sizeof(Item) * (size_t)C;
  if (bytes > f0->batch_bytes || C > f0->batch_cap) {
    cudaFree(f0->d_batch); cudaFree(f0->d_batch_sc);
    f0->d_batch = nullptr; f0->d_batch_sc = nullptr; f0->batch_bytes = 0; f0->batch_cap = 0;
    CU(cudaMalloc(&f0->d_batch, bytes));
    CU(cudaMalloc(&f0->d_batch_sc, sizeof(Scalars) * (size_t)C));
    f0->batch_bytes = bytes; f0->batch_cap = C;
  }
  std::vector<Item> items((size_t)C);
  for (int c = 0; c < C; ++c) {
    llpf_filter* f = fs[c];
    f->epoch = epochs ? epochs[c] : 0;
    Item& it = items[(size_t)c];
    std::memset(&it, 0, sizeof(it));
    base_params(f, it.P);
    if (it.P.nblocks != 1) return fail(LLPF_ERR_BAD_ARG, "llpf_run_batch needs filters created with single_block = 1");
    it.P.u = f0->d_u; it.P.y = f0->d_y;
    it.P.time_conv = is_aux(f) ? 0 : (conv == LLPF_TIME_LOGLIK ? 1 : 0);
    f->hsc = reset_scalars(f);      // the block performs reset! itself; stats_ahead etc. start from this state
    build_run_ops(f, it.P, (int)T, conv);
    fill_modelp<NX, NY>(f->hm, it.M);
    for (int r = 0; r < NX; ++r) {
      it.mu0[r] = f->hm.mu0[r];
      for (int cc = 0; cc <= r; ++cc) it.L0[r * MAX_NX + cc] = CMH(f->hm.L0, r, cc, NX);
    }
    it.sc0 = f->hsc;
  }
  CU(cudaMemcpyAsync(f0->d_batch, items.data(), bytes, cudaMemcpyHostToDevice, f0->stream));
  CU(cudaEventRecord(f0->ev0, f0->stream));
  void* d_items = f0->d_batch;
  Scalars* d_sc = f0->d_batch_sc;
  void* args[] = {(void*)&d_items, (void*)&d_sc};
  CU(cudaLaunchKernel(kernel, dim3(C), dim3(BLOCK), args, 0, f0->stream));
  CU(cudaEventRecord(f0->ev1, f0->stream));
  std::vector<Scalars> out((size_t)C);
  CU(cudaMemcpyAsync(out.data(), d_sc, sizeof(Scalars) * (size_t)C, cudaMemcpyDeviceToHost, f0->stream));
  CU(cudaStreamSynchronize(f0->stream));
  CU(cudaEventElapsedTime(&f0->last_ms, f0->ev0, f0->ev1));
  for (int c = 0; c < C; ++c) {
    fs[c]->hsc = out[(size_t)c];
    fs[c]->hsc.nonfinite = 0;       // a collapsed chain shows as a non-finite ll_out entry (PMMH maps it to -Inf)
    fs[c]->launches += 1;
    ll_out[c] = out[(size_t)c].ll_total;
  }
  return LLPF_OK;
}

extern "C" int llpf_run_batch(int32_t C, const llpf_handle* handles, int64_t T, const double* u, const double* y,
                              int32_t time_convention, const uint64_t* epochs, double* ll_out) {
  if (C < 1 || !handles || !y || !ll_out) return fail(LLPF_ERR_BAD_ARG, "bad argument");
  if (T < 1 || T > (1ll << 30)) return fail(LLPF_ERR_BAD_ARG, "bad T");
  llpf_filter* f0 = handles[0];
  OKR(check_handle(f0));
  if (f0->hm.nu > 0 && !u) return fail(LLPF_ERR_BAD_ARG, "u is null");
  for (int c = 0; c < C; ++c) {
    llpf_filter* f = handles[c];
    OKR(check_handle(f));
    if (f->wide || f->user || f->world > 1) return fail(LLPF_ERR_UNSUPPORTED, "llpf_run_batch: single-GPU descriptor-model Float64 filters");
    if (f->device != f0->device || f->hm.nx != f0->hm.nx || f->hm.ny != f0->hm.ny || f->hm.nu != f0->hm.nu ||
        f->hm.dyn != f0->hm.dyn || resid_of(f) != resid_of(f0) || f->cfg.filter != f0->cfg.filter)
      return fail(LLPF_ERR_BAD_ARG, "llpf_run_batch: all filters must share device, dimensions, dynamics kind, filter kind and resampling family");
    if (!f->cfg.single_block) return fail(LLPF_ERR_BAD_ARG, "llpf_run_batch needs filters created with single_block = 1");
  }
  const void* k = engine_batch_kernel(f0->hm.nx, f0->hm.ny, f0->hm.dyn, resid_of(f0));
  if (!k) return fail(LLPF_ERR_UNSUPPORTED, "no batched kernel instantiated for this (nx, ny, dynamics, resampling)");
  CU(cudaSetDevice(f0->device));
  OKR(ensure_run_buffers(f0, T));
  if (f0->hm.nu > 0)
    CU(cudaMemcpyAsync(f0->d_u, u, sizeof(double) * T * f0->hm.nu, cudaMemcpyHostToDevice, f0->stream));
  CU(cudaMemcpyAsync(f0->d_y, y, sizeof(double) * T * f0->hm.ny, cudaMemcpyHostToDevice, f0->stream));
  for (int c = 1; c < C; ++c) CU(cudaStreamSynchronize(handles[c]->stream));   // earlier work of the chains is done
  const int nx = f0->hm.nx, ny = f0->hm.ny;
  llpf_filter* const* fs = handles;
  if (nx == 1 && ny == 1) return run_batch_t<1, 1>(C, fs, T, time_convention, epochs, ll_out, k);
  if (nx == 2 && ny == 1) return run_batch_t<2, 1>(C, fs, T, time_convention, epochs, ll_out, k);
  if (nx == 2 && ny == 2) return run_batch_t<2, 2>(C, fs, T, time_convention, epochs, ll_out, k);
  if (nx == 4 && ny == 2) return run_batch_t<4, 2>(C, fs, T, time_convention, epochs, ll_out, k);
  return fail(LLPF_ERR_UNSUPPORTED, "no batched kernel instantiated for this (nx, ny)");
}

// ------------------------------------------------------------------------------------------------
// particle smoother (FFBS)  smoothing.jl:104-143
// ------------------------------------------------------------------------------------------------
struct Scratchpad {  // tiny RAII arena for the stand-alone calls
  std::vector<void*> ptrs;
  ~Scratchpad() { for (void* p : ptrs) cudaFree(p); }
  template <class T>
  cudaError_t alloc(T** p, size_t count) {
    cudaError_t e = cudaMalloc((void**)p, sizeof(T) * (count ? count : 1));
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

__global__ void k_fill_uniform53(double* out, long long n, RngKey key, uint32_t stream, uint32_t step) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint4 r = rng_block(key, stream, step, (unsigned long long)i, 0);
    out[i] = uniform53(r.x, r.y);
  }
}

// backward simulation on a forward history resident in device memory; xb_out is host memory [T][M][nx]
static int standalone_geometry(int device, long long n, EngineP& P, const void* kernel);
static void fixed_point_scale(const double* we, long long N, EngineP& P);

// weT_host: the caller's wef[:,T] when the history came from the host (llpf_smooth_history: weights of any scale), else
// nullptr (history written by the forward pass: normalised weights, 2^62 scale)
static int smooth_backward(llpf_filter* f, long long T, const double* u_dev, const DevHistory& H, long long M,
                           uint64_t epoch, double* xb_out, const double* weT_host = nullptr) {
  const long long N = f->N;
  const int nx = f->hm.nx;
  if (M < 1 || M > N) return fail(LLPF_ERR_BAD_ARG, "smooth: need 1 <= M <= N (smoothing.jl:122)");
  Scratchpad sp;
  double *d_us = nullptr, *d_xb = nullptr;
  long long* d_j = nullptr;
  CU(sp.alloc(&d_us, (size_t)M)); CU(sp.alloc(&d_j, (size_t)M)); CU(sp.alloc(&d_xb, (size_t)T * M * nx));
  CU(cudaMemsetAsync(d_j, 0, sizeof(long long) * M, f->stream));   // resample.jl:14: fresh j = zeros(Int, M)
  RngKey key{(uint32_t)f->cfg.seed, (uint32_t)(f->cfg.seed >> 32), (uint32_t)epoch << 8};
  k_fill_uniform53<<<(int)std::min<long long>((M + 255) / 256, 1024), 256, 0, f->stream>>>(d_us, M, key, ST_SMOOTH, 0u);
  CU(cudaGetLastError());
  // j = resample(pf.resampling_strategy, wef[:,T], M)   smoothing.jl:124 — the same kernels as the stand-alone entries,
  // on the filter's own scratch arrays and stream
  EngineP P;
  base_params(f, P);
  P.strategy = f->cfg.resampling;
  // block geometry from the occupancy of the kernel that is actually launched (not the engine's), scale from the data
  OKR(standalone_geometry(f->device, N, P, f->cfg.resampling == LLPF_RESAMPLE_RESIDUAL ? (const void*)k_resample_residual
                                                                                      : (const void*)k_resample));
  if (weT_host) fixed_point_scale(weT_host, N, P);
  const double* d_weT = H.we + (size_t)(T - 1) * N;
  int Mi = (int)M;
  CU(cudaMemsetAsync(f->bar, 0, sizeof(unsigned) * BAR_TOTAL_WORDS, f->stream));
  if (f->cfg.resampling == LLPF_RESAMPLE_RESIDUAL) {
    const double* d_us_c = d_us;
    void* args[] = {(void*)&P, (void*)&d_weT, (void*)&d_us_c, (void*)&Mi, (void*)&d_j};
    CU(cudaLaunchCooperativeKernel((const void*)k_resample_residual, dim3(P.nblocks), dim3(BLOCK), args, 0, f->stream));
  } else {
    double u01 = 0.0;
    const double* d_slots = nullptr;
    if (f->cfg.resampling == LLPF_RESAMPLE_SYSTEMATIC) {
      CU(cudaMemcpyAsync(&u01, d_us, sizeof(double), cudaMemcpyDeviceToHost, f->stream));
      CU(cudaStreamSynchronize(f->stream));
    } else {
      d_slots = d_us;
    }
    // (the weights live on the device: the literal range of resample.jl:24 is used, which is what a 53-bit rand() gives
    //  in all but ~1e-8 of the draws at smoother-sized N — llpf_julia_range.h)
    RangeArg range;
    std::memset(&range, 0, sizeof(range));
    void* args[] = {(void*)&P, (void*)&d_weT, (void*)&u01, (void*)&d_slots, (void*)&Mi, (void*)&d_j, (void*)&range};
    CU(cudaLaunchCooperativeKernel((const void*)k_resample, dim3(P.nblocks), dim3(BLOCK), args, 0, f->stream));
  }
  f->launches += 2;
  SmoothP S;
  std::memset(&S, 0, sizeof(S));
  S.xf = H.x; S.wf = H.w; S.u = u_dev; S.j0 = d_j; S.xb = d_xb;
  S.N = (int)N; S.M = (int)M; S.T = (int)T;
  S.Ts = f->cfg.Ts;
  double ld = 0.0;
  for (int i = 0; i < nx; ++i) ld += std::log(CMH(f->hm.L1, i, i, nx));
  S.c0 = -((double)nx * std::log(2 * M_PI) + 2 * ld) / 2;
  S.key = key;
  SmoothModelG G;
  std::memset(&G, 0, sizeof(G));
  std::vector<double> Winv;
  inv_lower(f->hm.L1, nx, Winv);
  for (int r = 0; r < nx; ++r) {
    for (int c = 0; c < nx; ++c) {
      if (!f->hm.A.empty()) G.A[r * MAX_NX + c] = CMH(f->hm.A, r, c, nx);
      G.Winv[r * MAX_NX + c] = CMH(Winv, r, c, nx);
    }
    for (int c = 0; c < f->hm.nu; ++c)
      if (!f->hm.B.empty()) G.B[r * MAX_NU + c] = CMH(f->hm.B, r, c, nx);
  }
  if (f->hm.dyn == LLPF_DYN_QUADTANK_RK4) {   // the quadtank coefficients, as fill_modelp packs them
    const HostModel& hq = f->hm;
    const double k1 = hq.dynp[1], k2 = hq.dynp[2], Aa = hq.dynp[3], a = hq.dynp[4], g = hq.dynp[5];
    G.qt[0] = -a / Aa; G.qt[1] = -(a * hq.a1_factor) / Aa; G.qt[2] = a / Aa; G.qt[3] = 2 * 9.81;
    G.qt[4] = g * k1 / Aa; G.qt[5] = g * k2 / Aa; G.qt[6] = (1 - g) * k2 / Aa; G.qt[7] = (1 - g) * k1 / Aa;
  }
  G.t_switch = f->hm.t_switch;
  G.integ_h = f->hm.integ_Ts / (double)f->hm.supersample;
  G.supersample = f->hm.supersample;
  G.nu = f->hm.nu;
  const int grid = (int)std::min<long long>(M, (long long)f->num_sms * 8);
  CU(cudaEventRecord(f->ev0, f->stream));
  CU(smooth_launch(nx, f->hm.dyn, S, G, grid, f->stream));
  CU(cudaEventRecord(f->ev1, f->stream));
  f->launches += 1;
  CU(cudaMemcpyAsync(xb_out, d_xb, sizeof(double) * (size_t)T * M * nx, cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  CU(cudaEventElapsedTime(&f->last_smooth_ms, f->ev0, f->ev1));
  return LLPF_OK;
}

static int smooth_supported(llpf_filter* f) {
  if (f->user) return fail(LLPF_ERR_UNSUPPORTED, "smooth: descriptor models only (the backward kernel has no user-function hook yet)");
  if (f->wide) return fail(LLPF_ERR_UNSUPPORTED, "smooth: Float32-particle filters keep no history");
  if (f->world > 1) return fail(LLPF_ERR_UNSUPPORTED, "smooth: single-GPU filters only (the history is not sharded)");
  return LLPF_OK;
}

extern "C" int llpf_smooth(llpf_handle h, int64_t T, const double* u, const double* y, int64_t M, uint64_t epoch,
                           double* ll, double* xb_out, const llpf_run_outputs* out) {
  OKR(check_handle(h));
  OKR(smooth_supported(h));
  if (!y || (h->hm.nu > 0 && !u) || !xb_out) return fail(LLPF_ERR_BAD_ARG, "u / y / xb_out is null");
  if (T < 1) return fail(LLPF_ERR_BAD_ARG, "bad T");
  if (M < 1 || M > h->N) return fail(LLPF_ERR_BAD_ARG, "smooth: need 1 <= M <= N (smoothing.jl:122)");
  CU(cudaSetDevice(h->device));
  OKR(ensure_run_buffers(h, T));
  if (h->hm.nu > 0)
    CU(cudaMemcpyAsync(h->d_u, u, sizeof(double) * T * h->hm.nu, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_y, y, sizeof(double) * T * h->hm.ny, cudaMemcpyHostToDevice, h->stream));
  DevHistory H;
  // sol = forward_trajectory(pf, u, y, p)   smoothing.jl:105 — the history stays in HBM
  int rc = run_impl(h, T, h->d_u, h->d_y, LLPF_TIME_FORWARD_TRAJECTORY, epoch, ll, out, &H);
  if (rc == LLPF_OK) rc = smooth_backward(h, T, h->d_u, H, M, epoch, xb_out);
  H.release();
  return rc;
}

extern "C" int llpf_smooth_history(llpf_handle h, int64_t T, const double* u, const double* xf, const double* wf,
                                   const double* wef, int64_t M, uint64_t epoch, double* xb_out) {
  OKR(check_handle(h));
  OKR(smooth_supported(h));
  if (!xf || !wf || !wef || (h->hm.nu > 0 && !u) || !xb_out) return fail(LLPF_ERR_BAD_ARG, "null argument");
  if (T < 1) return fail(LLPF_ERR_BAD_ARG, "bad T");
  if (M < 1 || M > h->N) return fail(LLPF_ERR_BAD_ARG, "smooth: need 1 <= M <= N (smoothing.jl:122)");
  CU(cudaSetDevice(h->device));
  OKR(ensure_run_buffers(h, T));
  if (h->hm.nu > 0)
    CU(cudaMemcpyAsync(h->d_u, u, sizeof(double) * T * h->hm.nu, cudaMemcpyHostToDevice, h->stream));
  DevHistory H;
  const size_t NT = (size_t)h->N * (size_t)T;
  int rc = LLPF_OK;
  cudaError_t e = cudaMalloc(&H.x, NT * h->hm.nx * 8);
  if (!e) e = cudaMalloc(&H.w, NT * 8);
  if (!e) e = cudaMalloc(&H.we, NT * 8);
  if (!e) e = cudaMemcpyAsync(H.x, xf, NT * h->hm.nx * 8, cudaMemcpyHostToDevice, h->stream);
  if (!e) e = cudaMemcpyAsync(H.w, wf, NT * 8, cudaMemcpyHostToDevice, h->stream);
  if (!e) e = cudaMemcpyAsync(H.we, wef, NT * 8, cudaMemcpyHostToDevice, h->stream);
  if (e) rc = fail(LLPF_ERR_CUDA, std::string("smooth: staging the history: ") + cudaGetErrorString(e));
  if (rc == LLPF_OK) rc = smooth_backward(h, T, h->d_u, H, M, epoch, xb_out, wef + (size_t)(T - 1) * h->N);
  H.release();
  return rc;
}


// ------------------------------------------------------------------------------------------------
// weighted statistics of the HBM-resident history (llpf_stats.cuh)
// ------------------------------------------------------------------------------------------------
typedef void (*moments_fn)(const double*, const double*, long long, int, double*, double*, double*, int, cudaStream_t);
template <int NX>
static void launch_moments(const double* xh, const double* weh, long long N, int T, double* m, double* mo, double* cv,
                           int grid, cudaStream_t st) {
  k_hist_moments<NX><<<grid, BLOCK, 0, st>>>(xh, weh, N, T, m, mo, cv);
}
static const moments_fn g_moments_by_nx[MAX_NX + 1] = {nullptr, launch_moments<1>, launch_moments<2>, launch_moments<3>,
                                                       launch_moments<4>, launch_moments<5>, launch_moments<6>,
                                                       launch_moments<7>, launch_moments<8>};

extern "C" int llpf_run_stats(llpf_handle h, int64_t T, const double* u, const double* y, uint64_t epoch, double* ll,
                              const llpf_run_outputs* out, const llpf_hist_stats* stats) {
  OKR(check_handle(h));
  if (!stats) return fail(LLPF_ERR_BAD_ARG, "stats is null");
  if (h->wide || h->world > 1) return fail(LLPF_ERR_UNSUPPORTED, "llpf_run_stats: single-GPU Float64-particle filters");
  if (!y || (h->hm.nu > 0 && !u)) return fail(LLPF_ERR_BAD_ARG, "u / y is null");
  if (T < 1) return fail(LLPF_ERR_BAD_ARG, "bad T");
  if (stats->nq < 0 || (stats->nq > 0 && (!stats->q || !stats->xquantile))) return fail(LLPF_ERR_BAD_ARG, "bad quantile request");
  for (int k = 0; k < stats->nq; ++k)
    if (!(stats->q[k] >= 0.0 && stats->q[k] <= 1.0)) return fail(LLPF_ERR_BAD_ARG, "quantile probabilities must lie in [0, 1]");
  CU(cudaSetDevice(h->device));
  OKR(ensure_run_buffers(h, T));
  if (h->hm.nu > 0)
    CU(cudaMemcpyAsync(h->d_u, u, sizeof(double) * T * h->hm.nu, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_y, y, sizeof(double) * T * h->hm.ny, cudaMemcpyHostToDevice, h->stream));
  DevHistory H;
  int rc = run_impl(h, T, h->d_u, h->d_y, LLPF_TIME_FORWARD_TRAJECTORY, epoch, ll, out, &H);
  if (rc != LLPF_OK) { H.release(); return rc; }
  const int nx = h->hm.nx, nq = stats->nq;
  Scratchpad sp;
  double *d_mean = nullptr, *d_mode = nullptr, *d_cov = nullptr, *d_q = nullptr, *d_qo = nullptr;
  cudaError_t e = cudaSuccess;
  if (stats->xmean) e = sp.alloc(&d_mean, (size_t)T * nx);
  if (!e && stats->xmode) e = sp.alloc(&d_mode, (size_t)T * nx);
  if (!e && stats->xcov) e = sp.alloc(&d_cov, (size_t)T * nx * nx);
  if (!e && nq > 0) e = sp.alloc(&d_q, (size_t)nq);
  if (!e && nq > 0) e = sp.alloc(&d_qo, (size_t)T * nq * nx);
  if (!e && nq > 0) e = cudaMemcpyAsync(d_q, stats->q, sizeof(double) * nq, cudaMemcpyHostToDevice, h->stream);
  if (!e && (d_mean || d_mode || d_cov)) {
    const int grid = (int)std::min<long long>(T, (long long)h->num_sms * 8);
    g_moments_by_nx[nx](H.x, H.we, h->N, (int)T, d_mean, d_mode, d_cov, grid, h->stream);
    e = cudaGetLastError();
    h->launches += 1;
  }
  if (!e && nq > 0) {
    const long long items = (long long)T * nx * nq;
    const int grid = (int)std::min<long long>(items, (long long)h->num_sms * 8);
    k_hist_quantile<<<grid, BLOCK, 0, h->stream>>>(H.x, H.we, h->N, (int)T, nx, d_q, nq, d_qo);
    e = cudaGetLastError();
    h->launches += 1;
  }
  if (!e && d_mean) e = cudaMemcpyAsync(stats->xmean, d_mean, sizeof(double) * T * nx, cudaMemcpyDeviceToHost, h->stream);
  if (!e && d_mode) e = cudaMemcpyAsync(stats->xmode, d_mode, sizeof(double) * T * nx, cudaMemcpyDeviceToHost, h->stream);
  if (!e && d_cov) e = cudaMemcpyAsync(stats->xcov, d_cov, sizeof(double) * T * nx * nx, cudaMemcpyDeviceToHost, h->stream);
  if (!e && d_qo) e = cudaMemcpyAsync(stats->xquantile, d_qo, sizeof(double) * T * nq * nx, cudaMemcpyDeviceToHost, h->stream);
  if (!e) e = cudaStreamSynchronize(h->stream);
  H.release();
  if (e) return fail(LLPF_ERR_CUDA, std::string("history statistics: ") + cudaGetErrorString(e));
  return LLPF_OK;
}

extern "C" int llpf_last_smooth_ms(llpf_handle h, float* ms) {
  OKR(check_handle(h));
  if (!ms) return fail(LLPF_ERR_BAD_ARG, "null");
  *ms = h->last_smooth_ms;
  return LLPF_OK;
}

// ------------------------------------------------------------------------------------------------
// accessors
// ------------------------------------------------------------------------------------------------
static int grid_for(long long n, int sms) {
  long long g = (n + 255) / 256;
  if (g > (long long)sms * 8) g = (long long)sms * 8;
  return g < 1 ? 1 : (int)g;
}

extern "C" int llpf_num_particles(llpf_handle h, int64_t* N) {
  OKR(check_handle(h));
  if (!N) return fail(LLPF_ERR_BAD_ARG, "null");
  *N = h->N;
  return LLPF_OK;
}
extern "C" int llpf_local_particles(llpf_handle h, int64_t* n, int64_t* first) {
  OKR(check_handle(h));
  if (n) *n = h->n;
  if (first) *first = h->first;
  return LLPF_OK;
}
extern "C" int llpf_index(llpf_handle h, int64_t* t) {
  OKR(check_handle(h));
  if (!t) return fail(LLPF_ERR_BAD_ARG, "null");
  *t = h->hsc.t_index;
  return LLPF_OK;
}
extern "C" int llpf_get_particles(llpf_handle h, double* x) {
  OKR(check_handle(h));
  if (!x) return fail(LLPF_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  if (h->wide)
    k_export_x_wide<<<grid_for(h->n, h->num_sms), 256, 0, h->stream>>>(reinterpret_cast<const float*>(h->x[h->hsc.cur]),
                                                                       h->n, h->hm.nx, h->scratch);
  else
    k_export_x<<<grid_for(h->n, h->num_sms), 256, 0, h->stream>>>(h->x[h->hsc.cur], h->ld, h->n, h->hm.nx, h->scratch);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(cudaMemcpyAsync(x, h->scratch, sizeof(double) * h->n * h->hm.nx, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LLPF_OK;
}
// at every API boundary xprev == x (copyto!(xprev,x) closes predict!, filtering.jl:151; reset! :8-9)
extern "C" int llpf_get_xprev(llpf_handle h, double* x) { return llpf_get_particles(h, x); }

static int materialise(llpf_filter* h, double* w_host, double* we_host) {
  CU(cudaSetDevice(h->device));
  double* dw = h->scratch;
  double* dwe = h->scratch + h->ld;
  k_materialise<<<grid_for(h->n, h->num_sms), 256, 0, h->stream>>>(h->w, h->sc, h->n, h->N, dw, dwe);
  CU(cudaGetLastError());
  h->launches += 1;
  if (w_host) CU(cudaMemcpyAsync(w_host, dw, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
  if (we_host) CU(cudaMemcpyAsync(we_host, dwe, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LLPF_OK;
}
extern "C" int llpf_get_weights(llpf_handle h, double* w) {
  OKR(check_handle(h));
  if (!w) return fail(LLPF_ERR_BAD_ARG, "null");
  return materialise(h, w, nullptr);
}
extern "C" int llpf_get_expweights(llpf_handle h, double* we) {
  OKR(check_handle(h));
  if (!we) return fail(LLPF_ERR_BAD_ARG, "null");
  return materialise(h, nullptr, we);
}
extern "C" int llpf_get_ancestors(llpf_handle h, int64_t* j) {
  OKR(check_handle(h));
  if (!j) return fail(LLPF_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  long long* dj = reinterpret_cast<long long*>(h->scratch);
  k_export_j<<<grid_for(h->n, h->num_sms), 256, 0, h->stream>>>(h->j, h->hsc.j_identity, h->n, h->first, dj);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(cudaMemcpyAsync(j, dj, sizeof(long long) * h->n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LLPF_OK;
}
extern "C" int llpf_get_bins(llpf_handle h, double* bins) {
  OKR(check_handle(h));
  if (!bins) return fail(LLPF_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(bins, h->bins, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return LLPF_OK;
}
extern "C" int llpf_set_state(llpf_handle h, const double* x, const double* w, int64_t t) {
  OKR(check_handle(h));
  if (!x || !w) return fail(LLPF_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->scratch, x, sizeof(double) * h->n * h->hm.nx, cudaMemcpyHostToDevice, h->stream));
  if (h->wide)
    k_import_x_wide<<<grid_for(h->n, h->num_sms), 256, 0, h->stream>>>(reinterpret_cast<float*>(h->x[h->hsc.cur]), h->n,
                                                                       h->hm.nx, h->scratch);
  else
    k_import_x<<<grid_for(h->n, h->num_sms), 256, 0, h->stream>>>(h->x[h->hsc.cur], h->ld, h->n, h->hm.nx, h->scratch);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(cudaMemcpyAsync(h->w, w, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
  h->hsc.uniform = 0; h->hsc.pend = 0; h->hsc.stats_ahead = 0; h->hsc.stats_valid = 0;
  h->hsc.j_identity = 1; h->hsc.t_index = t;
  return push_scalars(h);
}

// (sum we, sum we^2, sum we*x) on the materialised expweights, block partials finished on the host
static int weight_stats(llpf_filter* h, double* s, double* q, double* sx) {
  CU(cudaSetDevice(h->device));
  double* dwe = h->scratch + h->ld;
  k_materialise<<<grid_for(h->n, h->num_sms), 256, 0, h->stream>>>(h->w, h->sc, h->n, h->N, nullptr, dwe);
  const int grid = std::min(grid_for(h->n, h->num_sms), 256);
  const int PW = h->wide ? 2 + WNX : 2 + MAX_NX;
  if (h->wide)
    k_wstats_wide<<<grid, 256, 0, h->stream>>>(dwe, reinterpret_cast<const float*>(h->x[h->hsc.cur]), h->n, h->hm.nx,
                                               h->wstat);
  else
    k_wstats<<<grid, 256, 0, h->stream>>>(dwe, h->x[h->hsc.cur], h->ld, h->n, h->hm.nx, h->wstat);
  CU(cudaGetLastError());
  h->launches += 2;
  std::vector<double> part((size_t)grid * PW);
  CU(cudaMemcpyAsync(part.data(), h->wstat, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  double acc[2 + WNX] = {0};
  for (int b = 0; b < grid; ++b)
    for (int k = 0; k < PW; ++k) acc[k] += part[(size_t)b * PW + k];
  if (s) *s = acc[0];
  if (q) *q = acc[1];
  if (sx) for (int d = 0; d < h->hm.nx; ++d) sx[d] = acc[2 + d];
  return LLPF_OK;
}
// Sharded filters: the weights are normalised GLOBALLY, so ESS / weighted mean need the cross-rank reduction.  One
// reduce-only pass of the engine (a collective: every rank calls the accessor, like every other verb of a sharded
// filter) re-derives (max, sum e, sum e^2, sum e*x) over all ranks; the log-likelihood bookkeeping is left untouched.
static int refresh_stats_sharded(llpf_filter* h) {
  CU(cudaSetDevice(h->device));
  const Scalars before = h->hsc;
  EngineP P;
  base_params(h, P);
  P.u = h->stage_u; P.y = h->stage_y;
  P.use_t_override = 1; P.t_override = 0.0; P.want_xhat = 1;
  push_op(P, OP_PF, 0, 1, 1, 0, 0, OPF_SKIP_MEAS);
  OKR(launch(h, P, false));
  h->hsc.ll_last = before.ll_last; h->hsc.ll_total = before.ll_total; h->hsc.nonfinite = before.nonfinite;
  if (before.stats_ahead) {   // APF between predict! and correct!: w[] stays raw, correct! is still to be called
    h->hsc.pend = 0; h->hsc.stats_ahead = 1; h->hsc.stats_valid = before.stats_valid;
  }
  return push_scalars(h);
}

extern "C" int llpf_effective_particles(llpf_handle h, double* ess) {
  OKR(check_handle(h));
  if (!ess) return fail(LLPF_ERR_BAD_ARG, "null");
  if (h->world > 1) {
    OKR(refresh_stats_sharded(h));
    *ess = h->hsc.ess;
    return LLPF_OK;
  }
  double q = 0;
  OKR(weight_stats(h, nullptr, &q, nullptr));
  *ess = 1.0 / q;  // 1/sum(abs2, we)  resample.jl:2
  return LLPF_OK;
}
extern "C" int llpf_shouldresample(llpf_handle h, int32_t* yes) {
  OKR(check_handle(h));
  if (!yes) return fail(LLPF_ERR_BAD_ARG, "null");
  if (h->cfg.resample_threshold == 1.0) { *yes = 1; return LLPF_OK; }
  double ess = 0;
  OKR(llpf_effective_particles(h, &ess));
  *yes = ess < (double)h->N * h->cfg.resample_threshold ? 1 : 0;
  return LLPF_OK;
}
extern "C" int llpf_weighted_mean(llpf_handle h, double* xhat) {
  OKR(check_handle(h));
  if (!xhat) return fail(LLPF_ERR_BAD_ARG, "null");
  if (h->world > 1) {
    if (h->wide) return fail(LLPF_ERR_UNSUPPORTED, "weighted_mean of a sharded Float32-particle filter");
    OKR(refresh_stats_sharded(h));
    for (int d = 0; d < h->hm.nx; ++d) xhat[d] = h->hsc.xhat[d];
    return LLPF_OK;
  }
  double s = 0;
  OKR(weight_stats(h, &s, nullptr, xhat));
  // @assert sum(we) ≈ 1  filtering.jl:542
  if (!(std::fabs(s - 1.0) <= 1e-6)) return fail(LLPF_ERR_NONFINITE, "weights do not sum to one");
  return LLPF_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone numerics
// ------------------------------------------------------------------------------------------------

// power-of-two fixed-point scale so that sum(we)*scale stays below 2^62 (exact for normalised weights)
static void fixed_point_scale(const double* we, long long N, EngineP& P) {
  double sum = 0.0;
  for (long long i = 0; i < N; ++i) sum += (we[i] > 0 ? we[i] : 0.0);
  int e = 0;
  if (sum > 0 && std::isfinite(sum)) std::frexp(sum * (1.0 + 1e-9), &e);   // sum < 2^e
  if (e < 0) e = 0;
  P.fix_scale = std::ldexp(1.0, 62 - e);
  P.fix_inv = std::ldexp(1.0, e - 62);
}

static int standalone_geometry(int device, long long n, EngineP& P, const void* kernel) {
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  int occ = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, BLOCK, 0));
  if (occ < 1) return fail(LLPF_ERR_CUDA, "stand-alone kernel does not fit");
  long long nb = (n + BLOCK - 1) / BLOCK;
  const long long cap = std::min<long long>((long long)occ * prop.multiProcessorCount, MAX_BLOCKS - 1);
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  P.nblocks = (int)nb;
  P.chunk = (int)((n + nb - 1) / nb);
  P.chunk = (P.chunk + 1) & ~1;
  return LLPF_OK;
}

static int resample_standalone(int strategy, int64_t N, const double* we, double u01, const double* u_slots,
                               int64_t M, int64_t* j_inout, double* bins_out, int32_t scan_mode, int32_t device) {
  if (N < 1 || M < 1 || !we || !j_inout) return fail(LLPF_ERR_BAD_ARG, "bad argument");
  if (N >= (1ll << 31) || M >= (1ll << 31)) return fail(LLPF_ERR_BAD_ARG, "N, M < 2^31");
  int ndev = 0;
  OKR(llpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(LLPF_ERR_NO_DEVICE, "no such CUDA device");
  EngineP P;
  std::memset(&P, 0, sizeof(P));
  OKR(standalone_geometry(device, N, P, (const void*)k_resample));
  Scratchpad sp;
  double *d_we = nullptr, *d_bins = nullptr, *d_part = nullptr, *d_us = nullptr;
  u64 *d_tots = nullptr, *d_loc = nullptr;
  unsigned* d_bar = nullptr;
  long long* d_j = nullptr;
  CU(sp.alloc(&d_we, (size_t)N)); CU(sp.alloc(&d_bins, (size_t)N)); CU(sp.alloc(&d_part, (size_t)MAX_BLOCKS * PS));
  CU(sp.alloc(&d_tots, (size_t)MAX_BLOCKS)); CU(sp.alloc(&d_bar, BAR_TOTAL_WORDS)); CU(sp.alloc(&d_j, (size_t)M));
  CU(sp.alloc(&d_loc, (size_t)N));
  CU(cudaMemcpy(d_we, we, sizeof(double) * N, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_j, j_inout, sizeof(long long) * M, cudaMemcpyHostToDevice));
  CU(cudaMemset(d_bar, 0, BAR_TOTAL_WORDS * sizeof(unsigned)));
  if (u_slots) {
    CU(sp.alloc(&d_us, (size_t)M));
    CU(cudaMemcpy(d_us, u_slots, sizeof(double) * M, cudaMemcpyHostToDevice));
  }
  fixed_point_scale(we, N, P);
  // Julia's range constructor for the thresholds (resample.jl:24) needs r and bins[N] exactly as the kernel will find
  // them: SERIAL = the left-to-right f64 sum, FAST = the exact fixed-point total
  RangeArg range;
  std::memset(&range, 0, sizeof(range));
  if (strategy == LLPF_RESAMPLE_SYSTEMATIC) {
    double total;
    if (scan_mode != LLPF_SCAN_FAST) {
      volatile double acc = we[0];
      for (int64_t i = 1; i < N; ++i) acc = acc + we[i];
      total = acc;
    } else {
      unsigned long long tot = 0;
      for (int64_t i = 0; i < N; ++i) {
        const double v = we[i] * P.fix_scale;
        tot += (v > 0.0) ? (unsigned long long)std::nearbyint(std::fmin(v, FIX_SCALE)) : 0ull;
      }
      total = (double)tot * P.fix_inv;
    }
    volatile double ut = u01 * total;
    const double r = ut / (double)N;
    const RangeTP R = julia_range(r, 1.0 / (double)M, total + r);
    range.ref_hi = R.ref_hi; range.ref_lo = R.ref_lo; range.step_hi = R.step_hi; range.step_lo = R.step_lo;
    range.offset = R.offset; range.rational = R.rational;
  }
  int* d_heavy = nullptr;
  CU(sp.alloc(&d_heavy, (size_t)(1 + 3 * HEAVY_MAX)));
  CU(cudaMemset(d_heavy, 0, sizeof(int)));
  P.heavy = d_heavy;
  P.bins = d_bins; P.partials = d_part; P.tots = d_tots; P.bar = d_bar; P.loc = d_loc;
  P.N = N; P.n = (int)N; P.first = 0; P.strategy = strategy; P.scan_mode = scan_mode;
  P.world = 1;
  const double* d_us_c = d_us;
  int Mi = (int)M;
  void* args[] = {(void*)&P, (void*)&d_we, (void*)&u01, (void*)&d_us_c, (void*)&Mi, (void*)&d_j, (void*)&range};
  CU(cudaLaunchCooperativeKernel((const void*)k_resample, dim3(P.nblocks), dim3(BLOCK), args, 0, 0));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(j_inout, d_j, sizeof(long long) * M, cudaMemcpyDeviceToHost));
  if (bins_out) CU(cudaMemcpy(bins_out, d_bins, sizeof(double) * N, cudaMemcpyDeviceToHost));
  return LLPF_OK;
}

extern "C" int llpf_resample_systematic(int64_t N, const double* we, double u01, int64_t M, int64_t* j_inout,
                                        double* bins_out, int32_t scan_mode, int32_t device) {
  return resample_standalone(LLPF_RESAMPLE_SYSTEMATIC, N, we, u01, nullptr, M, j_inout, bins_out, scan_mode, device);
}
extern "C" int llpf_resample_stratified(int64_t N, const double* we, const double* u01, int64_t M,
                                        int64_t* j_inout, double* bins_out, int32_t scan_mode, int32_t device) {
  if (!u01) return fail(LLPF_ERR_BAD_ARG, "u01 is null");
  return resample_standalone(LLPF_RESAMPLE_STRATIFIED, N, we, 0.0, u01, M, j_inout, bins_out, scan_mode, device);
}

extern "C" int llpf_resample_residual(int64_t N, const double* we, const double* u01, int64_t M, int64_t* j_inout,
                                      double* bins_out, int32_t scan_mode, int32_t device) {
  if (N < 1 || M < 1 || !we || !j_inout || !u01) return fail(LLPF_ERR_BAD_ARG, "bad argument");
  if (N >= (1ll << 31) || M >= (1ll << 31)) return fail(LLPF_ERR_BAD_ARG, "N, M < 2^31");
  int ndev = 0;
  OKR(llpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(LLPF_ERR_NO_DEVICE, "no such CUDA device");
  EngineP P;
  std::memset(&P, 0, sizeof(P));
  OKR(standalone_geometry(device, N, P, (const void*)k_resample_residual));
  Scratchpad sp;
  double *d_we = nullptr, *d_bins = nullptr, *d_part = nullptr, *d_u = nullptr;
  u64 *d_tots = nullptr, *d_tots2 = nullptr, *d_loc = nullptr;
  unsigned* d_bar = nullptr;
  long long* d_j = nullptr;
  int* d_heavy = nullptr;
  CU(sp.alloc(&d_we, (size_t)N)); CU(sp.alloc(&d_bins, (size_t)N)); CU(sp.alloc(&d_part, (size_t)MAX_BLOCKS * PS));
  CU(sp.alloc(&d_tots, (size_t)MAX_BLOCKS)); CU(sp.alloc(&d_tots2, (size_t)MAX_BLOCKS));
  CU(sp.alloc(&d_bar, BAR_TOTAL_WORDS)); CU(sp.alloc(&d_j, (size_t)M)); CU(sp.alloc(&d_loc, (size_t)N));
  CU(sp.alloc(&d_u, (size_t)M)); CU(sp.alloc(&d_heavy, (size_t)(1 + 3 * HEAVY_MAX)));
  CU(cudaMemcpy(d_we, we, sizeof(double) * N, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_u, u01, sizeof(double) * M, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_j, j_inout, sizeof(long long) * M, cudaMemcpyHostToDevice));
  CU(cudaMemset(d_bar, 0, BAR_TOTAL_WORDS * sizeof(unsigned)));
  CU(cudaMemset(d_heavy, 0, sizeof(int)));
  fixed_point_scale(we, N, P);
  P.heavy = d_heavy;
  P.bins = d_bins; P.partials = d_part; P.tots = d_tots; P.tots2 = d_tots2; P.bar = d_bar; P.loc = d_loc;
  P.N = N; P.n = (int)N; P.first = 0; P.strategy = LLPF_RESAMPLE_RESIDUAL; P.scan_mode = scan_mode;
  P.world = 1;
  const double* d_u_c = d_u;
  int Mi = (int)M;
  void* args[] = {(void*)&P, (void*)&d_we, (void*)&d_u_c, (void*)&Mi, (void*)&d_j};
  CU(cudaLaunchCooperativeKernel((const void*)k_resample_residual, dim3(P.nblocks), dim3(BLOCK), args, 0, 0));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(j_inout, d_j, sizeof(long long) * M, cudaMemcpyDeviceToHost));
  if (bins_out) CU(cudaMemcpy(bins_out, d_bins, sizeof(double) * N, cudaMemcpyDeviceToHost));
  return LLPF_OK;
}


// Metropolis resampling at the function boundary (llpf_metropolis.cuh; extension, not in the reference)
__global__ void k_resample_metropolis(const double* __restrict__ we, long long N, long long M, int B, RngKey key,
                                      long long* __restrict__ j_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    long long k = i % N;
    double wk = we[k];
    for (int b = 0; 2 * b < B; ++b) {
      const uint4 r = rng_block(key, ST_METRO, 0u, (unsigned long long)i, (uint32_t)b);
      const long long j0 = (long long)__umulhi(r.x, (uint32_t)N), j1 = (long long)__umulhi(r.z, (uint32_t)N);
      const double w0 = we[j0], w1 = we[j1];
      if (uniform32_open(r.y) * wk <= w0) { k = j0; wk = w0; }
      if (2 * b + 1 < B && uniform32_open(r.w) * wk <= w1) { k = j1; wk = w1; }
    }
    j_out[i] = k + 1;
  }
}
extern "C" int llpf_resample_metropolis(int64_t N, const double* we, int64_t M, int32_t B, uint64_t seed, int64_t* j_out,
                                        int32_t device) {
  if (N < 1 || M < 1 || B < 1 || !we || !j_out) return fail(LLPF_ERR_BAD_ARG, "bad argument");
  if (N >= (1ll << 31) || M >= (1ll << 31)) return fail(LLPF_ERR_BAD_ARG, "N, M < 2^31");
  int ndev = 0;
  OKR(llpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(LLPF_ERR_NO_DEVICE, "no such CUDA device");
  CU(cudaSetDevice(device));
  Scratchpad sp;
  double* d_we = nullptr;
  long long* d_j = nullptr;
  CU(sp.alloc(&d_we, (size_t)N)); CU(sp.alloc(&d_j, (size_t)M));
  CU(cudaMemcpy(d_we, we, sizeof(double) * N, cudaMemcpyHostToDevice));
  RngKey key{(uint32_t)seed, (uint32_t)(seed >> 32), 0u};
  const int grid = (int)std::min<long long>((M + 255) / 256, 148 * 8);
  k_resample_metropolis<<<grid, 256>>>(d_we, N, M, B, key, d_j);
  CU(cudaGetLastError());
  CU(cudaMemcpy(j_out, d_j, sizeof(long long) * M, cudaMemcpyDeviceToHost));
  return LLPF_OK;
}

extern "C" int llpf_logsumexp(int64_t N, double* w, double* we, double* ll, int32_t device) {
  if (N < 1 || !w || !we || !ll) return fail(LLPF_ERR_BAD_ARG, "bad argument");
  int ndev = 0;
  OKR(llpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(LLPF_ERR_NO_DEVICE, "no such CUDA device");
  EngineP P;
  std::memset(&P, 0, sizeof(P));
  OKR(standalone_geometry(device, N, P, (const void*)k_logsumexp));
  Scratchpad sp;
  double *d_w = nullptr, *d_we = nullptr, *d_part = nullptr, *d_ll = nullptr;
  unsigned* d_bar = nullptr;
  CU(sp.alloc(&d_w, (size_t)N)); CU(sp.alloc(&d_we, (size_t)N)); CU(sp.alloc(&d_part, (size_t)MAX_BLOCKS * PS));
  CU(sp.alloc(&d_ll, 1)); CU(sp.alloc(&d_bar, BAR_TOTAL_WORDS));
  CU(cudaMemcpy(d_w, w, sizeof(double) * N, cudaMemcpyHostToDevice));
  CU(cudaMemset(d_bar, 0, BAR_TOTAL_WORDS * sizeof(unsigned)));
  P.partials = d_part; P.bar = d_bar; P.N = N; P.n = (int)N; P.world = 1;
  void* args[] = {(void*)&P, (void*)&d_w, (void*)&d_we, (void*)&d_ll};
  CU(cudaLaunchCooperativeKernel((const void*)k_logsumexp, dim3(P.nblocks), dim3(BLOCK), args, 0, 0));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(w, d_w, sizeof(double) * N, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(we, d_we, sizeof(double) * N, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ll, d_ll, sizeof(double), cudaMemcpyDeviceToHost));
  return LLPF_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: particles sharded over the GPUs of one box, exchange through IPC-mapped peer memory
// ------------------------------------------------------------------------------------------------
struct ShardBlob {
  unsigned long long magic;
  int rank, world, nx, pad;
  long long n, ld;
  unsigned long long arena_bytes, o_x0, o_x1, o_j, o_mbox, o_pack;
  cudaIpcMemHandle_t mem;
};
static const unsigned long long kBlobMagic = 0x4c4c504642323030ull;  // "LLPFB200"

extern "C" int llpf_shard_blob_size(size_t* bytes) {
  if (!bytes) return fail(LLPF_ERR_BAD_ARG, "null");
  *bytes = sizeof(ShardBlob);
  return LLPF_OK;
}
extern "C" int llpf_shard_export(llpf_handle h, void* blob) {
  OKR(check_handle(h));
  if (!blob) return fail(LLPF_ERR_BAD_ARG, "null");
  CU(cudaSetDevice(h->device));
  ShardBlob b;
  std::memset(&b, 0, sizeof(b));
  b.magic = kBlobMagic; b.rank = h->rank; b.world = h->world; b.nx = h->hm.nx;
  b.n = h->n; b.ld = h->ld; b.arena_bytes = h->arena_bytes;
  b.o_x0 = h->o_x0; b.o_x1 = h->o_x1; b.o_j = h->o_j; b.o_mbox = h->o_mbox; b.o_pack = h->o_pack;
  CU(cudaIpcGetMemHandle(&b.mem, h->arena));
  std::memcpy(blob, &b, sizeof(b));
  return LLPF_OK;
}
extern "C" int llpf_shard_connect(llpf_handle h, const void* blobs) {
  OKR(check_handle(h));
  if (!blobs) return fail(LLPF_ERR_BAD_ARG, "null");
  if (h->world <= 1) { h->connected = true; return LLPF_OK; }
  CU(cudaSetDevice(h->device));
  const ShardBlob* B = reinterpret_cast<const ShardBlob*>(blobs);
  for (int r = 0; r < h->world; ++r) {
    const ShardBlob& b = B[r];
    if (b.magic != kBlobMagic || b.rank != r || b.world != h->world || b.nx != h->hm.nx || b.n != h->n ||
        b.ld != h->ld || b.arena_bytes != h->arena_bytes || b.o_x0 != h->o_x0 || b.o_j != h->o_j || b.o_mbox != h->o_mbox ||
        b.o_pack != h->o_pack)
      return fail(LLPF_ERR_BAD_ARG, "shard blob mismatch (all ranks must create identical filters, blobs ordered by rank)");
  }
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank || h->peer_base[r]) continue;
    void* base = nullptr;
    CU(cudaIpcOpenMemHandle(&base, B[r].mem, cudaIpcMemLazyEnablePeerAccess));
    h->peer_base[r] = (char*)base;
  }
  h->connected = true;
  return LLPF_OK;
}

// ------------------------------------------------------------------------------------------------
// instrumentation
// ------------------------------------------------------------------------------------------------
extern "C" int llpf_launch_count(llpf_handle h, int64_t* launches) {
  OKR(check_handle(h));
  if (!launches) return fail(LLPF_ERR_BAD_ARG, "null");
  *launches = h->launches;
  return LLPF_OK;
}
extern "C" int llpf_last_run_ms(llpf_handle h, float* ms) {
  OKR(check_handle(h));
  if (!ms) return fail(LLPF_ERR_BAD_ARG, "null");
  *ms = h->last_ms;
  return LLPF_OK;
}
extern "C" int llpf_device_pointers(llpf_handle h, void** x_dev, void** w_dev, void** stream) {
  OKR(check_handle(h));
  if (x_dev) *x_dev = h->x[h->hsc.cur];
  if (w_dev) *w_dev = h->w;
  if (stream) *stream = (void*)h->stream;
  return LLPF_OK;
}

// ------------------------------------------------------------------------------------------------
// Ensemble Kalman filter (reference src/enkf.jl; kernel in llpf_enkf.cuh / llpf_enkf.cu)
// ------------------------------------------------------------------------------------------------
namespace llpf {
const void* enkf_kernel(int nx, int ny, int dyn);
// mirrors EnkfP of llpf_enkf.cuh (this translation unit does not include the kernel)
struct EnkfPHost {
  double* x; long long ld; long long N;
  int n, nblocks, chunk, nu;
  unsigned int* bar; double* partials;
  const double* u; const double* y;
  int T, do_correct, do_predict, use_t_single;
  double Ts, t_single, inflation;
  RngKey key;
  double* st;
  double *o_x, *o_R, *o_xt, *o_Rt, *o_e, *o_ll, *o_S, *o_K;
};
constexpr int kEnkfKmax = 128;   // ENKF_KMAX
}  // namespace llpf

static int enkf_prepare(llpf_filter* f, const void** kernel) {
  if (f->hm.wide || f->hm.dyn == LLPF_DYN_USER || f->world > 1 || f->cfg.filter != LLPF_FILTER_PF)
    return fail(LLPF_ERR_UNSUPPORTED, "EnKF verbs need a single-GPU Float64 ParticleFilter-kind handle with descriptor dynamics "
                                      "and a linear measurement");
  *kernel = enkf_kernel(f->hm.nx, f->hm.ny, f->hm.dyn);
  if (!*kernel) return fail(LLPF_ERR_UNSUPPORTED, "no EnKF kernel instantiated for this (nx, ny, dynamics)");
  const int nx = f->hm.nx;
  if (!f->enkf_st) {
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, *kernel, BLOCK, 0));
    if (occ < 1) return fail(LLPF_ERR_CUDA, "EnKF kernel does not fit on an SM");
    long long nb = (f->n + BLOCK - 1) / BLOCK;
    nb = std::min<long long>(nb, std::min(occ * f->num_sms, MAX_BLOCKS - 1));
    f->enkf_blocks = (int)std::max<long long>(nb, 1);
    CU(cudaMalloc(&f->enkf_st, sizeof(double) * (size_t)(2 + nx + nx * nx + 1)));
    CU(cudaMemsetAsync(f->enkf_st, 0, sizeof(double) * (size_t)(2 + nx + nx * nx + 1), f->stream));
    CU(cudaMalloc(&f->enkf_partials, sizeof(double) * (size_t)2 * f->enkf_blocks * kEnkfKmax));
  }
  return LLPF_OK;
}

// one launch of k_enkf; `out` (host, nullable pointers) receives the per-step outputs of the T steps
struct EnkfHostOut { double *x, *R, *xt, *Rt, *e, *ll, *S, *K; };
static int enkf_launch(llpf_filter* f, int T, const double* d_u, const double* d_y, int do_correct, int do_predict,
                       bool use_t, double t, const EnkfHostOut& out) {
  const void* k = nullptr;
  OKR(enkf_prepare(f, &k));
  const int nx = f->hm.nx, ny = f->hm.ny;
  const size_t Tn = (size_t)std::max(T, 1);
  const size_t sizes[8] = {Tn * nx, Tn * nx * nx, Tn * nx, Tn * nx * nx, Tn * ny, Tn, Tn * ny * ny, Tn * nx * ny};
  double* const host[8] = {out.x, out.R, out.xt, out.Rt, out.e, out.ll, out.S, out.K};
  size_t need = 0;
  for (int q = 0; q < 8; ++q) if (host[q]) need += sizes[q];
  if (need > f->enkf_out_doubles) {
    cudaFree(f->enkf_out); f->enkf_out = nullptr; f->enkf_out_doubles = 0;
    CU(cudaMalloc(&f->enkf_out, sizeof(double) * need));
    f->enkf_out_doubles = need;
  }
  double* dev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  { size_t o = 0; for (int q = 0; q < 8; ++q) if (host[q]) { dev[q] = f->enkf_out + o; o += sizes[q]; } }
  llpf::EnkfPHost P;
  std::memset(&P, 0, sizeof(P));
  P.x = f->x[f->hsc.cur]; P.ld = f->ld; P.N = f->N; P.n = (int)f->n;
  P.nblocks = f->enkf_blocks;
  P.chunk = (int)((f->n + P.nblocks - 1) / P.nblocks);
  P.nu = f->hm.nu;
  P.bar = f->bar; P.partials = f->enkf_partials;
  P.u = d_u; P.y = d_y;
  P.T = T; P.do_correct = do_correct; P.do_predict = do_predict; P.use_t_single = use_t ? 1 : 0;
  P.Ts = f->cfg.Ts; P.t_single = t; P.inflation = f->enkf_inflation;
  P.key = RngKey{(uint32_t)f->cfg.seed, (uint32_t)(f->cfg.seed >> 32), (uint32_t)f->epoch << 8};
  P.st = f->enkf_st;
  P.o_x = dev[0]; P.o_R = dev[1]; P.o_xt = dev[2]; P.o_Rt = dev[3]; P.o_e = dev[4]; P.o_ll = dev[5]; P.o_S = dev[6]; P.o_K = dev[7];
  std::vector<double> M = pack_modelp_runtime(f->hm);
  std::vector<double> E((size_t)ny * nx + 2 * (size_t)ny * ny, 0.0);   // EnkfM: C, R2, L2 row-major
  for (int a = 0; a < ny; ++a) {
    for (int c = 0; c < nx; ++c) E[(size_t)a * nx + c] = CMH(f->hm.C, a, c, ny);
    for (int c = 0; c < ny; ++c) {
      E[(size_t)ny * nx + (size_t)a * ny + c] = CMH(f->hm.R2, a, c, ny);
      E[(size_t)ny * nx + (size_t)ny * ny + (size_t)a * ny + c] = CMH(f->hm.L2, a, c, ny);
    }
  }
  CU(cudaMemsetAsync(f->bar, 0, sizeof(unsigned) * BAR_TOTAL_WORDS, f->stream));
  void* args[] = {(void*)&P, (void*)M.data(), (void*)E.data()};
  CU(cudaEventRecord(f->ev0, f->stream));
  CU(cudaLaunchCooperativeKernel(k, dim3(P.nblocks), dim3(BLOCK), args, 0, f->stream));
  CU(cudaEventRecord(f->ev1, f->stream));
  f->launches += 1;
  for (int q = 0; q < 8; ++q)
    if (host[q]) CU(cudaMemcpyAsync(host[q], dev[q], sizeof(double) * sizes[q], cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  CU(cudaEventElapsedTime(&f->last_ms, f->ev0, f->ev1));
  return LLPF_OK;
}

static int enkf_read_state(llpf_filter* f, double* ll, double* tindex, double* mean, double* cov) {
  const int nx = f->hm.nx;
  std::vector<double> st((size_t)2 + nx + nx * nx + 1);
  CU(cudaMemcpyAsync(st.data(), f->enkf_st, sizeof(double) * st.size(), cudaMemcpyDeviceToHost, f->stream));
  CU(cudaStreamSynchronize(f->stream));
  if (ll) *ll = st[0];
  if (tindex) *tindex = st[1];
  if (mean) std::memcpy(mean, st.data() + 2, sizeof(double) * nx);
  if (cov) std::memcpy(cov, st.data() + 2 + nx, sizeof(double) * nx * nx);
  if (st.back() != 0.0) return fail(LLPF_ERR_NOT_POSDEF, "Cholesky factorization of innovation covariance failed (enkf.jl:324)");
  if (!(std::fabs(st[0]) <= 1.7976931348623157e308)) return fail(LLPF_ERR_NONFINITE, "EnKF log-likelihood is not finite");
  return LLPF_OK;
}

extern "C" int llpf_enkf_set_inflation(llpf_handle h, double inflation) {
  OKR(check_handle(h));
  if (!(inflation > 0.0)) return fail(LLPF_ERR_BAD_ARG, "inflation must be positive");
  h->enkf_inflation = inflation;
  return LLPF_OK;
}

// reset!(enkf)  enkf.jl:205-224: redraw the ensemble from d0 (RNG stream 0 of `epoch`), t = 0, cached mean / covariance
extern "C" int llpf_enkf_reset(llpf_handle h, uint64_t epoch) {
  OKR(check_handle(h));
  const void* k = nullptr;
  OKR(enkf_prepare(h, &k));
  OKR(llpf_reset(h, epoch));
  const int nx = h->hm.nx;
  CU(cudaMemsetAsync(h->enkf_st, 0, sizeof(double) * (size_t)(2 + nx + nx * nx + 1), h->stream));
  return enkf_launch(h, 0, nullptr, nullptr, 0, 0, true, 0.0, EnkfHostOut{});
}

// state(enkf), covariance(enkf)  enkf.jl:186-193 (cached by the last verb); cov row-major nx*nx (symmetric); t = enkf.t
extern "C" int llpf_enkf_state(llpf_handle h, double* mean, double* cov, int64_t* t_index) {
  OKR(check_handle(h));
  if (!h->enkf_st) return fail(LLPF_ERR_BAD_ARG, "call llpf_enkf_reset first");
  CU(cudaSetDevice(h->device));
  double ti = 0;
  std::vector<double> st((size_t)2 + h->hm.nx + h->hm.nx * h->hm.nx + 1);
  CU(cudaMemcpyAsync(st.data(), h->enkf_st, sizeof(double) * st.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  ti = st[1];
  if (mean) std::memcpy(mean, st.data() + 2, sizeof(double) * h->hm.nx);
  if (cov) std::memcpy(cov, st.data() + 2 + h->hm.nx, sizeof(double) * h->hm.nx * h->hm.nx);
  if (t_index) *t_index = (int64_t)ti;
  return LLPF_OK;
}

// predict!(enkf, u, p, t)  enkf.jl:228-272
extern "C" int llpf_enkf_predict(llpf_handle h, const double* u, double t) {
  OKR(check_handle(h));
  if (!h->enkf_st) return fail(LLPF_ERR_BAD_ARG, "call llpf_enkf_reset first");
  CU(cudaSetDevice(h->device));
  OKR(stage_inputs(h, u, nullptr, nullptr));
  OKR(enkf_launch(h, 1, h->stage_u, h->stage_y, 0, 1, true, t, EnkfHostOut{}));
  return enkf_read_state(h, nullptr, nullptr, nullptr, nullptr);
}

// correct!(enkf, u, y, p, t) -> (; ll, e, S, K)  enkf.jl:281-356 (S, K row-major ny*ny / nx*ny; any of them may be NULL)
extern "C" int llpf_enkf_correct(llpf_handle h, const double* u, const double* y, double t, double* ll, double* e, double* S,
                                 double* K) {
  OKR(check_handle(h));
  if (!h->enkf_st) return fail(LLPF_ERR_BAD_ARG, "call llpf_enkf_reset first");
  if (!y) return fail(LLPF_ERR_BAD_ARG, "y is null");
  CU(cudaSetDevice(h->device));
  OKR(stage_inputs(h, u, y, nullptr));
  EnkfHostOut out{};
  out.e = e; out.S = S; out.K = K; out.ll = ll;
  OKR(enkf_launch(h, 1, h->stage_u, h->stage_y, 1, 0, true, t, out));
  return enkf_read_state(h, nullptr, nullptr, nullptr, nullptr);
}

// forward_trajectory(enkf, u, y)  filtering.jl:282-325: reset!, then per step record (x, R), correct!, record (xt, Rt, e),
// predict!; one launch.  Outputs (host, any may be NULL): x, xt [T][nx]; R, Rt [T][nx*nx]; e [T][ny]; ll_steps [T];
// S [T][ny*ny]; K [T][nx*ny].  *ll = sum of the step log-likelihoods.
extern "C" int llpf_enkf_run(llpf_handle h, int64_t T, const double* u, const double* y, uint64_t epoch, double* ll,
                             double* x, double* R, double* xt, double* Rt, double* e, double* ll_steps, double* S, double* K) {
  OKR(check_handle(h));
  if (T < 1 || T > (1 << 30) || !y) return fail(LLPF_ERR_BAD_ARG, "need T >= 1 and y");
  const int nu = h->hm.nu, ny = h->hm.ny;
  if (nu > 0 && !u) return fail(LLPF_ERR_BAD_ARG, "u is null");
  OKR(llpf_enkf_reset(h, epoch));
  OKR(ensure_run_buffers(h, T));
  if (nu > 0) CU(cudaMemcpyAsync(h->d_u, u, sizeof(double) * (size_t)T * nu, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_y, y, sizeof(double) * (size_t)T * ny, cudaMemcpyHostToDevice, h->stream));
  EnkfHostOut out{};
  out.x = x; out.R = R; out.xt = xt; out.Rt = Rt; out.e = e; out.ll = ll_steps; out.S = S; out.K = K;
  OKR(enkf_launch(h, (int)T, h->d_u, h->d_y, 1, 1, false, 0.0, out));
  return enkf_read_state(h, ll, nullptr, nullptr, nullptr);
}
