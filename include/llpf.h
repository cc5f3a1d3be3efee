/*
 * llpf.h — C-ABI of libllpf_b200.so: the B200-native particle-filter hot path.
 *
 * This is the drop-in boundary for ONE path of baggepinnen/LowLevelParticleFilters.jl:
 * the per-timestep correct!/predict!/update! loop + resample of ParticleFilter,
 * AdvancedParticleFilter and AuxiliaryParticleFilter, and the two trajectory drivers
 * forward_trajectory / loglik that run it.  The reference has no FFI layer (pure Julia,
 * multiple dispatch); each entry point below names the reference method it replaces
 * (file:line under the reference tree).  A Julia maintainer binds these with `ccall`
 * (see INTEGRATION.md and julia/LLPFB200.jl); the Python host mirror binds them with ctypes.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns an int status (LLPF_OK == 0).
 *  - matrices are COLUMN-MAJOR (Julia `Matrix{Float64}` memory order), leading dim = #rows.
 *  - u is nu x T, y is ny x T column-major == T contiguous vectors (Julia Vector{SVector} layout).
 *  - particle export is AoS: nx contiguous doubles per particle (Vector{SVector{nx,Float64}}),
 *    regardless of the SoA layout used in HBM.
 *  - indices returned to the caller (ancestors j) are 1-based Int64, like the reference.
 *  - user closures cannot cross a C-ABI: models are descriptors (llpf_model).
 *  - one handle = one CUDA stream = one mutable filter state (like one PFstate); not thread-safe.
 *  - host pointers unless the name says `_dev`.
 */
#ifndef LLPF_H
#define LLPF_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLPF_VERSION 200

/* ---- status codes -------------------------------------------------------------------- */
enum {
  LLPF_OK = 0,
  LLPF_ERR_BAD_ARG = 1,       /* null pointer / bad size / unsupported combination        */
  LLPF_ERR_CUDA = 2,          /* a CUDA runtime call failed (llpf_last_error has the text) */
  LLPF_ERR_NONFINITE = 3,     /* weight collapse: log-likelihood became non-finite         */
  LLPF_ERR_NO_DEVICE = 4,     /* no CUDA device / driver: the product path never falls back to CPU */
  LLPF_ERR_UNSUPPORTED = 5,   /* model dimension / option not compiled in                  */
  LLPF_ERR_NOT_POSDEF = 6     /* a covariance matrix failed Cholesky                       */
};

/* ---- filter kinds (src/PFtypes.jl:21-49, :162-177) ------------------------------------ */
enum {
  LLPF_FILTER_PF = 0,           /* ParticleFilter            PFtypes.jl:21-36   */
  LLPF_FILTER_ADVANCED = 1,     /* AdvancedParticleFilter    PFtypes.jl:162-177 */
  LLPF_FILTER_AUX = 2,          /* AuxiliaryParticleFilter{ParticleFilter}  PFtypes.jl:38-49, filtering.jl:195-217 */
  LLPF_FILTER_AUX_ADVANCED = 3  /* AuxiliaryParticleFilter{AdvancedParticleFilter}  filtering.jl:219-234 */
};

/* ---- resampling strategies (src/LowLevelParticleFilters.jl:43-46, src/resample.jl) ---- */
enum {
  LLPF_RESAMPLE_SYSTEMATIC = 0, /* resample.jl:17-36  */
  LLPF_RESAMPLE_STRATIFIED = 1, /* resample.jl:38-61  */
  LLPF_RESAMPLE_RESIDUAL = 2,   /* resample.jl:63-117 (single-GPU filters)           */
  LLPF_RESAMPLE_METROPOLIS = 3  /* NOT in the reference: Metropolis resampling (Murray, Lee & Jacob 2016), the low-
                                   synchronisation alternative: no prefix sum, ratios of weights only, biased for a
                                   finite number of proposals (llpf_config.metropolis_steps).  Single-GPU filters;
                                   (nx, ny) in {(4,2), (2,1), (2,2)} or any user-defined model.  Statistical tests only. */
};

/* ---- how the cumulative sum `bins` is formed (resample.jl:19-22) ----------------------- */
enum {
  LLPF_SCAN_FAST = 0,   /* device-wide parallel prefix sum (re-associated; differs from the
                           serial cumsum by O(sqrt(N)) ulp; bit-identical whenever all partial
                           sums are exactly representable, e.g. dyadic weights)              */
  LLPF_SCAN_SERIAL = 1  /* strict left-to-right f64 adds, the reference's order: bit-exact
                           `bins` and therefore bit-exact indices; verification mode (slow)  */
};

/* ---- particle element type (the eltype of `initial_density`, PFtypes.jl:66,202) ---------- */
enum {
  LLPF_PARTICLE_F64 = 0,  /* Vector{SVector{nx,Float64}}: nx <= 8, SoA in HBM, all filter kinds            */
  LLPF_PARTICLE_F32 = 1   /* Vector{SVector{nx,Float32}}: nx, ny <= 64, linear-Gaussian, ParticleFilter /
                             AdvancedParticleFilter (test/test_large.jl regime).  Weights, bins and the
                             log-likelihood stay Float64 like the reference (PFtypes.jl:68-69); the C-ABI
                             still exchanges particles as doubles                                          */
};

/* ---- dynamics descriptors --------------------------------------------------------------- */
enum {
  LLPF_DYN_LINEAR = 0,        /* x+ = A x + B u                 examples/example_lineargaussian.jl:27 */
  LLPF_DYN_QUADTANK_RK4 = 1,  /* quadruple-tank ODE, RK4        examples/example_quadtank.jl:91-106, src/utils.jl:220-237 */
  LLPF_DYN_USER = 2           /* dynamics / measurement log-likelihood given as CUDA device source (llpf_create_user):
                                 the reference's arbitrary closures `dynamics(x,u,p,t)` (PFtypes.jl:128,255) and
                                 `measurement_likelihood(x,u,y,p,t)` (PFtypes.jl:232)                                   */
};

/* time convention of the trajectory drivers:
 *   forward_trajectory passes t = (k-1)*Ts   (filtering.jl:352)
 *   loglik / callable filters pass t = index*Ts with index starting at 1 after reset! (filtering.jl:13,181,238) */
enum {
  LLPF_TIME_FORWARD_TRAJECTORY = 0,
  LLPF_TIME_LOGLIK = 1
};

/*
 * Model descriptor. Gaussian additive noise everywhere (the reference's MvNormal / SimpleMvNormal,
 * src/utils.jl:241-273): dynamics noise N(0,R1), measurement noise N(0,R2), initial state N(mu0,Sigma0).
 * Measurement is linear: y = C x  (quadtank: C = [I2 0], example_quadtank.jl:33).
 * For LLPF_DYN_QUADTANK_RK4, dyn_params = {kc,k1,k2,A,a,gamma} (example_quadtank.jl:91-97,109),
 * t_switch/a1_factor implement the hard-coded variant's `if t > 500; a1 *= 2` (:15-17); set
 * t_switch = +inf to disable.  integ_Ts / supersample are the rk4 arguments (utils.jl:220).
 */
typedef struct llpf_model {
  int32_t nx, nu, ny;
  int32_t dynamics;           /* LLPF_DYN_* */
  const double* A;            /* nx*nx  (LINEAR)            */
  const double* B;            /* nx*nu  (LINEAR)            */
  const double* C;            /* ny*nx                      */
  const double* R1;           /* nx*nx  dynamics-noise cov  */
  const double* R2;           /* ny*ny  measurement cov     */
  const double* mu0;          /* nx                         */
  const double* Sigma0;       /* nx*nx                      */
  double dyn_params[8];
  double t_switch;
  double a1_factor;
  double integ_Ts;
  int32_t supersample;
  int32_t _pad;
} llpf_model;

/*
 * Filter configuration: the keyword arguments of the reference constructors
 * (PFtypes.jl:21-36: resample_threshold=0.1, resampling_strategy=ResampleSystematic, rng, Ts=1.0;
 *  PFtypes.jl:162-177: resample_threshold=0.5).
 * `seed` replaces `rng`: noise is a counter-based Philox4x32-10 stream keyed by
 * (seed, epoch; stream, step, particle) — see DESIGN.md "RNG contract".
 */
typedef struct llpf_config {
  int64_t N;                  /* number of particles                                   */
  int32_t filter;             /* LLPF_FILTER_*                                         */
  int32_t resampling;         /* LLPF_RESAMPLE_* (RESIDUAL: world == 1 only)            */
  double  resample_threshold; /* resample.jl:5-10 ; ==1 means always                   */
  double  Ts;                 /* sample time                                           */
  uint64_t seed;
  int32_t scan_mode;          /* LLPF_SCAN_*                                           */
  int32_t device;             /* CUDA device ordinal                                   */
  /* particle sharding across GPUs of one box: this handle owns global particle
     indices [rank*N/world, (rank+1)*N/world). world==1 for a single-GPU filter.         */
  int32_t rank;
  int32_t world;
  int32_t particle_dtype;     /* LLPF_PARTICLE_*                                        */
  int32_t single_block;       /* 1: the whole filter is run by ONE thread block (small N: PMMH-sized filters); required
                                 for llpf_run_batch, and makes a chain's result independent of how it is launched     */
  int32_t metropolis_steps;   /* LLPF_RESAMPLE_METROPOLIS: proposals per output slot (0 = default 32)                  */
  int32_t _reserved;          /* must be 0                                                                             */
} llpf_config;

/* Optional outputs of llpf_run. Any pointer may be NULL. Host memory. */
typedef struct llpf_run_outputs {
  double* ll_steps;        /* [T]      per-step log-likelihood increments (correct! return value) */
  double* ess_steps;       /* [T]      effective_particles after each correct!  (resample.jl:1-2)  */
  int32_t* resampled;      /* [T]      1 if predict! resampled at that step                        */
  double* xhat;            /* [nx*T]   weighted_mean after each correct! (filtering.jl:541-568)    */
  /* full history, the x/w/we fields of ParticleFilteringSolution (solutions.jl:334-345),
     column-major N x T like the reference: x_hist is [T][N][nx] AoS, w/we are [T][N]             */
  double* x_hist;
  double* w_hist;
  double* we_hist;
} llpf_run_outputs;

/* Weighted statistics of the forward history, reduced on the device (the N x T history stays in HBM; only T x (...) results
   are copied back).  Any output pointer may be NULL.  Host memory.
     xmean     [T][nx]      mean_trajectory(sol)  = sum(x .* we)            filtering.jl:417,436-438
     xmode     [T][nx]      mode_trajectory(sol)  = particle with the largest weight (first on ties)   :427,434
     xcov      [T][nx][nx]  weighted_cov(sol): StatsBase cov, ProbabilityWeights, corrected = true     :575-583
     xquantile [T][nq][nx]  weighted_quantile(sol, q[k]): StatsBase quantile with ProbabilityWeights   :592-595   */
typedef struct llpf_hist_stats {
  double* xmean;
  double* xmode;
  double* xcov;
  const double* q;         /* [nq] probabilities in [0, 1] */
  int32_t nq;
  int32_t _pad;
  double* xquantile;
} llpf_hist_stats;

typedef struct llpf_filter* llpf_handle;

/* ---- life cycle ------------------------------------------------------------------------- */
/* constructors ParticleFilter(N, dynamics, measurement, df, dg, d0; kw...) PFtypes.jl:65-75,
   AdvancedParticleFilter(...) :200-210, AuxiliaryParticleFilter(...) :38-49.
   Allocates all device state (x, xprev, w, we, j, bins: PFstate, PFtypes.jl:8-17) and draws the
   initial particles (epoch 0).                                                              */
int llpf_create(const llpf_config* cfg, const llpf_model* model, llpf_handle* out);
int llpf_destroy(llpf_handle h);
const char* llpf_last_error(void);
int llpf_device_count(int* count);

/* replace model matrices / parameters (`p` overridden per call, filtering.jl:140,164) */
int llpf_set_model(llpf_handle h, const llpf_model* model);

/* User-defined models: the reference takes arbitrary Julia closures for `dynamics` and `measurement_likelihood`
   (PFtypes.jl:122-139, :226-239, :242-289).  Closures cannot cross a C-ABI; device source can.  `cuda_source` must define
       namespace llpf_user {
         template <> __device__ void   dynamics<NX>(double (&x)[NX], const double* u, const double* p, double t);   // x <- f(x,u,p,t), no noise
         template <> __device__ double loglik<NX>(const double (&x)[NX], const double* u, const double* y, const double* p, double t);
       }
   for NX == model->nx.  It is compiled at run time (NVRTC, sm_100a) together with the engine, so both functions are
   inlined into the fused sweep of k_engine<nx, ny, LLPF_DYN_USER>; every filter kind, resampling strategy, the step
   verbs, the trajectory drivers and sharding work as for the descriptor models.  Any 1 <= nx, ny <= 8.
   `model`: nx, nu, ny, R1 (the additive dynamics noise N(0,R1) is drawn by the engine, PFtypes.jl:135), mu0, Sigma0 are
   used; model->dynamics must be LLPF_DYN_USER; A, B, C, R2 may be NULL.  p[np]: the parameter vector `p` handed to both
   functions (copied to the device; llpf_set_user_params replaces it, e.g. between PMMH proposals, without recompiling).
   Compile errors: LLPF_ERR_BAD_ARG with the NVRTC log in llpf_last_error().  Needs libnvrtc.so.12 at run time
   (searched in the loader path, $LLPF_NVRTC_PATH, /usr/local/cuda/lib64).
   Particles with per-particle sufficient statistics (the reference's RBParticle = nonlinear state + mean and covariance
   of a per-particle Kalman filter, rbpf.jl:1-5, :163-283): a source that contains the token LLPF_USER_STATE_HOOKS must
   also define
         template <> __device__ void add_noise<NX>(double (&x)[NX], const double (&xprev)[NX], const double (&nz)[NX],
                                                   const double* u, const double* p, double t);   // replaces x += nz
         template <> __device__ void correct_state<NX>(double (&x)[NX], const double* u, const double* y,
                                                       const double* p, double t);   // mutation of the particle by correct!
   (ParticleFilter / AdvancedParticleFilter kinds; R1 and Sigma0 may then be positive SEMI-definite: deterministic
   components have zero rows).  The Python / Julia `RBPF` constructors generate such a source from the matrices.        */
int llpf_create_user(const llpf_config* cfg, const llpf_model* model, const char* cuda_source,
                     const double* p, int32_t np, llpf_handle* out);
int llpf_set_user_params(llpf_handle h, const double* p, int32_t np);

/* reset!(pf)  filtering.jl:4-14 : x=xprev ~ initial_density, w=-log N, we=1/N, t=1.
   `epoch` selects the RNG sub-stream (successive reset! calls in the reference advance pf.rng). */
int llpf_reset(llpf_handle h, uint64_t epoch);

/* ---- step verbs --------------------------------------------------------------------------- */
/* correct!(pf,u,y,p,t) -> (ll,0)  filtering.jl:164-168 (PF/Advanced), :170-174 (APF: logsumexp only).
   y containing NaN plays the role of `missing` (PFtypes.jl:109,227): the weight update is skipped. */
int llpf_correct(llpf_handle h, const double* u, const double* y, double t, double* ll);
/* predict!(pf,u,p,t)  filtering.jl:140-153 */
int llpf_predict(llpf_handle h, const double* u, double t);
/* predict!(pfa,u,y1,p,t)  filtering.jl:195-217 (and :219-234 for AUX_ADVANCED) */
int llpf_predict_aux(llpf_handle h, const double* u, const double* y1, double t);
/* update!(pf,u,y,p,t)  filtering.jl:181-185 ; APF: update!(pfa,u,y,y1,p,t) :187-191 (y1 may be NULL otherwise) */
int llpf_update(llpf_handle h, const double* u, const double* y, const double* y1, double t, double* ll);

/* ---- trajectory drivers (the fused device loop) ------------------------------------------- */
/* forward_trajectory(pf,u,y,p)  filtering.jl:343-365 (PF/Advanced), :367-384 (APF)
   loglik(pf,u,y,p)              smoothing.jl:227-230,            :232-236 (APF)
   Runs reset!(epoch) then all T steps on the device without returning to the host.
   time_convention picks which of the two the call reproduces (they differ in t and, for the APF,
   in the last step).                                                                          */
int llpf_run(llpf_handle h, int64_t T, const double* u, const double* y,
             int32_t time_convention, uint64_t epoch, double* ll, const llpf_run_outputs* out);
/* same, with u (nu*T) and y (ny*T) already resident in device memory; nothing is copied H2D */
int llpf_run_dev(llpf_handle h, int64_t T, const double* u_dev, const double* y_dev,
                 int32_t time_convention, uint64_t epoch, double* ll, const llpf_run_outputs* out);

/* forward_trajectory(pf,u,y,p) (filtering.jl:343-384) whose N x T history of x / w / we is kept in device memory, reduced
   there to the statistics requested in `stats` and then released: mean_trajectory / mode_trajectory / weighted_cov /
   weighted_quantile of the solution without moving the history to the host.  `out` as in llpf_run
   (x_hist / w_hist / we_hist may still be requested).  Single-GPU Float64-particle filters.                        */
int llpf_run_stats(llpf_handle h, int64_t T, const double* u, const double* y, uint64_t epoch, double* ll,
                   const llpf_run_outputs* out, const llpf_hist_stats* stats);

/* Batched multi-chain loglik: C independent filters — each its own handle (model, seed, state),
   created with cfg.single_block = 1 and identical dimensions / dynamics kind / resampling strategy, on one device —
   evaluated by ONE kernel launch, one thread block per filter: reset!(epochs[c]) + the T fused steps of
   loglik / forward_trajectory (time_convention) on the shared data u (nu*T), y (ny*T).  ll_out[c] = the log-likelihood of
   chain c, bit-identical to llpf_run on the same handle.  This is what `metropolis_threaded` (smoothing.jl:335-347)
   needs: one launch per MCMC iteration for all chains.                                                          */
int llpf_run_batch(int32_t C, const llpf_handle* handles, int64_t T, const double* u, const double* y,
                   int32_t time_convention, const uint64_t* epochs, double* ll_out);

/* ---- particle smoother: forward filtering, backward simulation -------------- */
/* xb, ll = smooth(pf, M, u, y, p)  smoothing.jl:104-107 : forward_trajectory (the N x T history of x, w, we stays in
   device memory) followed by M backward-simulation trajectories (smoothing.jl:116-143, draw_one_categorical
   resample.jl:128-152).  xb_out: [T][M][nx] doubles (the reference's M x T Matrix{SVector}, column-major).
   `out` (may be NULL) receives the forward pass's optional outputs as in llpf_run.  Needs 1 <= M <= N (:122).
   rand() draws come from the (seed, epoch) counter streams: DESIGN.md "RNG contract", stream 6.
   Single-GPU, Float64-particle filters.                                                                      */
int llpf_smooth(llpf_handle h, int64_t T, const double* u, const double* y, int64_t M, uint64_t epoch,
                double* ll, double* xb_out, const llpf_run_outputs* out);
/* xb, ll = smooth(pf, xf, wf, wef, ll, M, u, y, p)  smoothing.jl:116-143 with a caller-provided forward history:
   xf [T][N][nx], wf / wef [T][N] (ParticleFilteringSolution x / w / we, host memory)                        */
int llpf_smooth_history(llpf_handle h, int64_t T, const double* u, const double* xf, const double* wf,
                        const double* wef, int64_t M, uint64_t epoch, double* xb_out);
/* device time of the last backward-simulation kernel (ms, CUDA events on the handle's stream) */
int llpf_last_smooth_ms(llpf_handle h, float* ms);

/* ---- accessors (PFtypes.jl:296-334) ------------------------------------------------------- */
int llpf_num_particles(llpf_handle h, int64_t* N);      /* global N                           */
int llpf_local_particles(llpf_handle h, int64_t* n, int64_t* first); /* this shard's slice    */
int llpf_index(llpf_handle h, int64_t* t);              /* index(pf) = state.t[]              */
int llpf_get_particles(llpf_handle h, double* x);       /* particles(pf): [n][nx] AoS         */
int llpf_get_xprev(llpf_handle h, double* x);           /* state.xprev                        */
int llpf_get_weights(llpf_handle h, double* w);         /* weights(pf)    : log-weights       */
int llpf_get_expweights(llpf_handle h, double* we);     /* expweights(pf)                     */
int llpf_get_ancestors(llpf_handle h, int64_t* j);      /* state.j, 1-based global indices    */
int llpf_get_bins(llpf_handle h, double* bins);         /* state.bins of the last resample    */
int llpf_set_state(llpf_handle h, const double* x, const double* w, int64_t t); /* PFstate re-wrap, PFtypes.jl:77-81 */
int llpf_effective_particles(llpf_handle h, double* ess);  /* resample.jl:1-2  */
int llpf_shouldresample(llpf_handle h, int32_t* yes);      /* resample.jl:5-10 */
int llpf_weighted_mean(llpf_handle h, double* xhat);       /* filtering.jl:541-548,567-568 */

/* ---- stand-alone numerics at the reference's function boundaries ------------------------- */
/* resample(ResampleSystematic, we, j, bins, M)  resample.jl:17-36 with the rand() at :23 supplied as u01.
   we: [N] host; j_inout: [M] host Int64 1-based, pre-filled by the caller (entries with s[i] >= bins[N]
   keep their previous value, exactly like the reference); bins_out: [N] or NULL.               */
int llpf_resample_systematic(int64_t N, const double* we, double u01, int64_t M,
                             int64_t* j_inout, double* bins_out, int32_t scan_mode, int32_t device);
/* resample(ResampleStratified, ...) resample.jl:38-61 with the M rand() draws supplied as u01[M] */
int llpf_resample_stratified(int64_t N, const double* we, const double* u01, int64_t M,
                             int64_t* j_inout, double* bins_out, int32_t scan_mode, int32_t device);
/* resample(ResampleResidual, ...) resample.jl:63-117 with the rand() draws of :106 supplied in draw order as
   u01[M] (only the first M - sum(floor(we_i/sum(we)*M)) are consumed).  bins_out receives the reference's `bins`
   after the call: the normalised cumulative residuals (or the raw residuals when no draw was needed, :85-87).
   LLPF_SCAN_SERIAL performs the three sums (:66-69, :89-92, :99-102) left to right like the reference.        */
int llpf_resample_residual(int64_t N, const double* we, const double* u01, int64_t M,
                           int64_t* j_inout, double* bins_out, int32_t scan_mode, int32_t device);
/* Metropolis resampling at the function boundary (extension, see LLPF_RESAMPLE_METROPOLIS): j_out[M] 1-based ancestors,
   slot i starts its chain at particle i mod N; B proposals per slot; variates from (seed, stream 7).               */
int llpf_resample_metropolis(int64_t N, const double* we, int64_t M, int32_t B, uint64_t seed, int64_t* j_out,
                             int32_t device);
/* logsumexp!(w, we) utils.jl:18-27 : in-place on host arrays w[N] (normalised log-weights out),
   we[N] out, returns ll = log(sum(exp(w_in)))                                                  */
int llpf_logsumexp(int64_t N, double* w, double* we, double* ll, int32_t device);

/* ---- multi-GPU (one process per GPU) ------------------------------------------- */
/* Each rank creates its handle with cfg.rank/world set, then exchanges the opaque IPC blob
   (llpf_shard_blob_size bytes) with every other rank (torch.distributed / NCCL all_gather on the
   host side) and passes the concatenation [world][blob] to llpf_shard_connect.  After that the
   kernels exchange the per-step (max, sum exp, sum exp^2) partials and pull resampled particles
   directly over NVLink peer memory; no host round trip inside the time loop.                    */
int llpf_shard_blob_size(size_t* bytes);
int llpf_shard_export(llpf_handle h, void* blob);
int llpf_shard_connect(llpf_handle h, const void* blobs_all_ranks);

/* ---- instrumentation ------------------------------------------------------------------------ */
/* number of kernel launches issued by this handle since creation, and the device time of the last
   llpf_run* call measured with CUDA events on the handle's stream (ms)                          */
int llpf_launch_count(llpf_handle h, int64_t* launches);
int llpf_last_run_ms(llpf_handle h, float* ms);
/* raw device pointers of the SoA state for zero-copy consumers: x is [nx][n] doubles, w is [n]  */
int llpf_device_pointers(llpf_handle h, void** x_dev, void** w_dev, void** stream);

/* ---- Ensemble Kalman filter (reference src/enkf.jl: stochastic EnKF with perturbed observations) --------------------
 * SURVEY §8f rank 4.  The ensemble is the particle buffer of an ordinary handle: create it with llpf_create (filter =
 * LLPF_FILTER_PF, Float64 particles, descriptor dynamics, world = 1; C is the linear measurement, R1 / R2 / mu0 / Sigma0 as in
 * EnsembleKalmanFilter(dynamics, measurement, R1, R2, d0, N), enkf.jl:94-141), then use these verbs instead of the
 * particle-filter ones.  Matrices in the outputs are row-major (symmetric ones either way).                              */
int llpf_enkf_set_inflation(llpf_handle h, double inflation);                      /* kwarg `inflation`  enkf.jl:106,261-266 */
int llpf_enkf_reset(llpf_handle h, uint64_t epoch);                                /* reset!(enkf)       enkf.jl:205-224     */
int llpf_enkf_state(llpf_handle h, double* mean, double* cov, int64_t* t_index);   /* state / covariance / index :178-193    */
int llpf_enkf_predict(llpf_handle h, const double* u, double t);                   /* predict!           enkf.jl:228-272     */
int llpf_enkf_correct(llpf_handle h, const double* u, const double* y, double t,   /* correct! -> (; ll, e, S, K) :281-356   */
                      double* ll, double* e, double* S, double* K);
/* forward_trajectory(enkf, u, y) (filtering.jl:282-325) in ONE launch: reset!(epoch), then per step (x, R) recorded,
   correct!, (xt, Rt, e) recorded, predict!.  Host outputs, any may be NULL: x, xt [T][nx]; R, Rt [T][nx*nx]; e [T][ny];
   ll_steps [T]; S [T][ny*ny]; K [T][nx*ny]; *ll = sum of ll_steps.                                                          */
int llpf_enkf_run(llpf_handle h, int64_t T, const double* u, const double* y, uint64_t epoch, double* ll, double* x,
                  double* R, double* xt, double* Rt, double* e, double* ll_steps, double* S, double* K);

#ifdef __cplusplus
}
#endif
#endif /* LLPF_H */
