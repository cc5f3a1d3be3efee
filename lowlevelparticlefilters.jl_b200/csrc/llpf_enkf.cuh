// llpf_enkf.cuh — Ensemble Kalman filter (stochastic EnKF with perturbed observations), reference src/enkf.jl.
//
// SURVEY §8f rank 4 ("the other ensemble-parallel filter").  Same persistent cooperative design as the particle-filter
// engine: one launch runs a verb (reset statistics / predict! / correct!) or a whole forward_trajectory; the ensemble lives
// in the handle's SoA particle buffer; the dynamics, the noise L1*z and the counter-based RNG are the engine's own device
// functions (llpf_engine.cuh).  Per time step (forward_trajectory(kf::AbstractKalmanFilter), filtering.jl:282-325):
//   correct!  enkf.jl:281-356   C1  y_i = C x_i, sum x, sum y                      -> xbar, ybar
//                               C2  sum Ya Ya', sum Xa Ya'                          -> S = YaYa'/(N-1) + R2 (symmetrized), chol,
//                                                                                    K = (XaYa'/(N-1)) / chol(S), e = y - ybar, ll
//                               C3  x_i += K (y + eps_i - y_i), eps_i = L2 z_i ; sum x  -> enkf.x
//                               C4  sum (x_i - xbar)(x_i - xbar)'                  -> enkf.R      (_update_ensemble_stats!, :172-176)
//   predict!  enkf.jl:228-272   P1  x_i = f(x_i,u,p,t) + L1 z_i ; sum x            -> xbar
//                               P1b (inflation > 1) x_i = xbar + inflation (x_i - xbar) ; sum x
//                               P2  covariance                                      -> enkf.x, enkf.R ; t += 1
// Every sum is a deterministic grid reduction (fixed order inside a block, block partials added in block order by every
// block: all blocks hold bit-identical statistics and compute S, K redundantly).  The reference accumulates sequentially
// (`_ensemble_mean`, `mul!(R, dx, dx', 1, 1)`), so sums agree to rounding, not bit for bit: parity tolerance 1e-9.
// Measurement: linear, y = C x (descriptor); dynamics: the engine's descriptors (linear, quadtank RK4).
// RNG streams: initial ensemble 0, process noise 1 (step = enkf.t), observation perturbations 8 (step = enkf.t).
#pragma once
#include "llpf_engine.cuh"

namespace llpf {

constexpr uint32_t ST_ENKF_OBS = 8;
constexpr int ENKF_KMAX = 128;      // widest reduction: ny*ny + nx*ny with nx, ny <= 8

template <int NX, int NY>
struct EnkfM {
  double C[NY * NX];     // row-major
  double R2[NY * NY];
  double L2[NY * NY];    // lower Cholesky factor of R2 (perturbed observations eps = L2 z)
};

struct EnkfP {
  double* x;             // [nx][ld] SoA ensemble
  long long ld;
  long long N;
  int n, nblocks, chunk, nu;
  unsigned int* bar;     // grid barrier counter (zeroed by the host before the launch)
  double* partials;      // [2][nblocks][ENKF_KMAX]
  const double* u;       // [T][nu]
  const double* y;       // [T][ny]
  int T;
  int do_correct, do_predict;
  int use_t_single;      // verbs: t given by the caller; trajectory: t = (k-1)*Ts (filtering.jl:282 `t = range(0, step=Ts, ...)`)
  double Ts, t_single, inflation;
  RngKey key;
  double* st;            // [0] ll of the launch, [1] enkf.t, [2..2+nx) enkf.x, then enkf.R (nx*nx row-major), then status flag
  double *o_x, *o_R, *o_xt, *o_Rt, *o_e, *o_ll, *o_S, *o_K;   // per-step outputs of a trajectory / the verb (device, nullable)
};

template <int K>
__device__ __forceinline__ void enkf_grid_sum(const EnkfP& P, double (&v)[K], double* red, double* tot, unsigned& bar_target,
                                              unsigned& seq) {
  static_assert(K <= ENKF_KMAX, "reduction too wide");
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  __syncthreads();                                   // previous users of red / tot are done
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) red[(threadIdx.x >> 5) * ENKF_KMAX + k] = v[k];
  }
  __syncthreads();
  double* part = P.partials + ((size_t)(seq & 1u) * P.nblocks + LLPF_BLOCKIDX) * ENKF_KMAX;
  for (int k = threadIdx.x; k < K; k += BLOCK) {
    double s = red[k];
    for (int w = 1; w < NWARP; ++w) s += red[w * ENKF_KMAX + k];
    __stcg(part + k, s);
  }
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  const double* all = P.partials + (size_t)(seq & 1u) * P.nblocks * ENKF_KMAX;
  for (int k = threadIdx.x; k < K; k += BLOCK) {
    double s = __ldcg(all + k);
    for (int b = 1; b < P.nblocks; ++b) s += __ldcg(all + (size_t)b * ENKF_KMAX + k);
    tot[k] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = tot[k];
  seq += 1;
}

template <int NX, int NY, int DYN>
__global__ void __launch_bounds__(BLOCK, 1)
k_enkf(const __grid_constant__ EnkfP P, const __grid_constant__ ModelP<NX, NY> M, const __grid_constant__ EnkfM<NX, NY> E) {
  __shared__ Shared sh;
  __shared__ double red[NWARP * ENKF_KMAX];
  __shared__ double tot[ENKF_KMAX];
  math_tab_load(sh.mt);
  model_to_shared<NX, NY>(M, sh);
  __syncthreads();
  unsigned bar_target = 0, seq = 0;
  int beg, end;
  {
    long long b = (long long)blockIdx.x * P.chunk, e = b + P.chunk;
    if (b > P.n) b = P.n;
    if (e > P.n) e = P.n;
    beg = (int)b; end = (int)e;
  }
  const double invN = 1.0 / (double)P.N, invN1 = 1.0 / (double)(P.N - 1);
  double ll_total = 0.0;
  int t_index = (int)__ldcg(P.st + 1);
  int status = 0;
  double mean[NX], cov[NX * NX];
#pragma unroll
  for (int r = 0; r < NX; ++r) mean[r] = __ldcg(P.st + 2 + r);
#pragma unroll
  for (int k = 0; k < NX * NX; ++k) cov[k] = __ldcg(P.st + 2 + NX + k);
  const bool out0 = (blockIdx.x == 0 && threadIdx.x == 0);

  auto measure = [&](const double (&x)[NX], double (&yv)[NY]) {
#pragma unroll
    for (int a = 0; a < NY; ++a) {
      double acc = E.C[a * NX] * x[0];
#pragma unroll
      for (int c = 1; c < NX; ++c) acc = fma(E.C[a * NX + c], x[c], acc);
      yv[a] = acc;
    }
  };
  // _update_ensemble_stats!: enkf.x = mean(ensemble) (its sum is already in `s`), enkf.R = sum (x - xbar)(x - xbar)' / (N - 1)
  auto covariance_pass = [&](double (&s)[NX]) {
#pragma unroll
    for (int r = 0; r < NX; ++r) mean[r] = s[r] * invN;
    double c2[NX * NX];
#pragma unroll
    for (int k = 0; k < NX * NX; ++k) c2[k] = 0.0;
    for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
      double d[NX];
#pragma unroll
      for (int r = 0; r < NX; ++r) d[r] = __ldcg(P.x + (size_t)r * P.ld + i) - mean[r];
#pragma unroll
      for (int r = 0; r < NX; ++r)
#pragma unroll
        for (int c = 0; c < NX; ++c) c2[r * NX + c] = fma(d[r], d[c], c2[r * NX + c]);
    }
    enkf_grid_sum<NX * NX>(P, c2, red, tot, bar_target, seq);
#pragma unroll
    for (int k = 0; k < NX * NX; ++k) cov[k] = c2[k] * invN1;
  };
  auto put_state = [&](double* ox, double* oR, int k) {
    if (!out0) return;
    if (ox) {
#pragma unroll
      for (int r = 0; r < NX; ++r) ox[(size_t)k * NX + r] = mean[r];
    }
    if (oR) {
#pragma unroll
      for (int q = 0; q < NX * NX; ++q) oR[(size_t)k * NX * NX + q] = cov[q];
    }
  };

  if (!P.do_correct && !P.do_predict) {   // statistics of the ensemble as it is (reset!, set_state)
    double s[NX];
#pragma unroll
    for (int r = 0; r < NX; ++r) s[r] = 0.0;
    for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
#pragma unroll
      for (int r = 0; r < NX; ++r) s[r] += __ldcg(P.x + (size_t)r * P.ld + i);
    }
    enkf_grid_sum<NX>(P, s, red, tot, bar_target, seq);
    covariance_pass(s);
  }

  for (int k = 0; k < P.T; ++k) {
    const double t = P.use_t_single ? P.t_single : (double)k * P.Ts;
    if (P.do_correct) {
      put_state(P.o_x, P.o_R, k);                    // x[k], R[k]: the prediction  filtering.jl:297-298
      double yobs[NY];
#pragma unroll
      for (int a = 0; a < NY; ++a) yobs[a] = __ldg(P.y + (size_t)k * NY + a);
      // C1
      double s1[NX + NY];
#pragma unroll
      for (int q = 0; q < NX + NY; ++q) s1[q] = 0.0;
      for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
        double x[NX], yv[NY];
        load_x<NX>(P.x, P.ld, i, x);
        measure(x, yv);
#pragma unroll
        for (int r = 0; r < NX; ++r) s1[r] += x[r];
#pragma unroll
        for (int a = 0; a < NY; ++a) s1[NX + a] += yv[a];
      }
      enkf_grid_sum<NX + NY>(P, s1, red, tot, bar_target, seq);
      double xb[NX], yb[NY];
#pragma unroll
      for (int r = 0; r < NX; ++r) xb[r] = s1[r] * invN;
#pragma unroll
      for (int a = 0; a < NY; ++a) yb[a] = s1[NX + a] * invN;
      // C2
      double s2[NY * NY + NX * NY];
#pragma unroll
      for (int q = 0; q < NY * NY + NX * NY; ++q) s2[q] = 0.0;
      for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
        double x[NX], yv[NY];
        load_x<NX>(P.x, P.ld, i, x);
        measure(x, yv);
#pragma unroll
        for (int a = 0; a < NY; ++a) yv[a] -= yb[a];
#pragma unroll
        for (int a = 0; a < NY; ++a)
#pragma unroll
          for (int b = 0; b < NY; ++b) s2[a * NY + b] = fma(yv[a], yv[b], s2[a * NY + b]);
#pragma unroll
        for (int r = 0; r < NX; ++r) {
          const double d = x[r] - xb[r];
#pragma unroll
          for (int a = 0; a < NY; ++a) s2[NY * NY + r * NY + a] = fma(d, yv[a], s2[NY * NY + r * NY + a]);
        }
      }
      enkf_grid_sum<NY * NY + NX * NY>(P, s2, red, tot, bar_target, seq);
      // S = symmetrize(Ya Ya' / (N-1) + R2) ; chol ; K = (Xa Ya' / (N-1)) / chol(S) ; e ; ll      enkf.jl:318-350
      double S0[NY * NY], S[NY * NY], Ls[NY * NY], Kg[NX * NY], ev[NY];
#pragma unroll
      for (int q = 0; q < NY * NY; ++q) S0[q] = s2[q] * invN1 + E.R2[q];
#pragma unroll
      for (int a = 0; a < NY; ++a)
#pragma unroll
        for (int b = 0; b < NY; ++b) S[a * NY + b] = 0.5 * (S0[a * NY + b] + S0[b * NY + a]);
#pragma unroll
      for (int q = 0; q < NY * NY; ++q) Ls[q] = 0.0;
#pragma unroll
      for (int jc = 0; jc < NY; ++jc) {
        double d = S[jc * NY + jc];
#pragma unroll
        for (int m = 0; m < jc; ++m) d -= Ls[jc * NY + m] * Ls[jc * NY + m];
        if (!(d > 0.0)) status = 1;               // "Cholesky factorization of innovation covariance failed"  enkf.jl:324
        d = sqrt(d);
        Ls[jc * NY + jc] = d;
#pragma unroll
        for (int i2 = jc + 1; i2 < NY; ++i2) {
          double v = S[i2 * NY + jc];
#pragma unroll
          for (int m = 0; m < jc; ++m) v -= Ls[i2 * NY + m] * Ls[jc * NY + m];
          Ls[i2 * NY + jc] = v / d;
        }
      }
#pragma unroll
      for (int r = 0; r < NX; ++r) {
        double row[NY], yv2[NY];
#pragma unroll
        for (int a = 0; a < NY; ++a) row[a] = s2[NY * NY + r * NY + a] * invN1;
#pragma unroll
        for (int a = 0; a < NY; ++a) {
          double v = row[a];
#pragma unroll
          for (int m = 0; m < a; ++m) v -= Ls[a * NY + m] * yv2[m];
          yv2[a] = v / Ls[a * NY + a];
        }
#pragma unroll
        for (int a = NY - 1; a >= 0; --a) {
          double v = yv2[a];
#pragma unroll
          for (int m = a + 1; m < NY; ++m) v -= Ls[m * NY + a] * Kg[r * NY + m];
          Kg[r * NY + a] = v / Ls[a * NY + a];
        }
      }
      double q2 = 0.0, ld = 0.0;
      {
        double w2[NY];
#pragma unroll
        for (int a = 0; a < NY; ++a) {
          ev[a] = yobs[a] - yb[a];
          double v = ev[a];
#pragma unroll
          for (int m = 0; m < a; ++m) v -= Ls[a * NY + m] * w2[m];
          w2[a] = v / Ls[a * NY + a];
          q2 += w2[a] * w2[a];
          ld += log(Ls[a * NY + a]);
        }
      }
      const double ll = -(NY * 1.8378770664093453 + 2.0 * ld) / 2 - q2 / 2;   // extended_logpdf  utils.jl:252-257
      ll_total += ll;
      // C3: perturbed-observation update
      double s3[NX];
#pragma unroll
      for (int r = 0; r < NX; ++r) s3[r] = 0.0;
      for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
        double x[NX], yv[NY], z[NY];
        load_x<NX>(P.x, P.ld, i, x);
        measure(x, yv);
        normals<NY>(P.key, ST_ENKF_OBS, (uint32_t)t_index, (unsigned long long)(unsigned)i, z, sh.mt);
        double dy[NY];
#pragma unroll
        for (int a = 0; a < NY; ++a) {
          double eps = E.L2[a * NY] * z[0];
#pragma unroll
          for (int m = 1; m <= a; ++m) eps = fma(E.L2[a * NY + m], z[m], eps);
          dy[a] = (yobs[a] + eps) - yv[a];           // yi_pert .- yi_pred   enkf.jl:341-343
        }
#pragma unroll
        for (int r = 0; r < NX; ++r) {
          double acc = Kg[r * NY] * dy[0];
#pragma unroll
          for (int a = 1; a < NY; ++a) acc = fma(Kg[r * NY + a], dy[a], acc);
          x[r] += acc;
          s3[r] += x[r];
        }
        store_x<NX>(P.x, P.ld, i, x);
      }
      enkf_grid_sum<NX>(P, s3, red, tot, bar_target, seq);
      covariance_pass(s3);                           // C4
      put_state(P.o_xt, P.o_Rt, k);                  // xt[k], Rt[k]  filtering.jl:305-306
      if (out0) {
        if (P.o_ll) P.o_ll[k] = ll;
        if (P.o_e) {
#pragma unroll
          for (int a = 0; a < NY; ++a) P.o_e[(size_t)k * NY + a] = ev[a];
        }
        if (P.o_S) {
#pragma unroll
          for (int q = 0; q < NY * NY; ++q) P.o_S[(size_t)k * NY * NY + q] = S[q];
        }
        if (P.o_K) {
#pragma unroll
          for (int q = 0; q < NX * NY; ++q) P.o_K[(size_t)k * NX * NY + q] = Kg[q];
        }
      }
    }
    if (P.do_predict) {
      __syncthreads();
      if (threadIdx.x < NX) {
        const double* u = P.u + (size_t)k * P.nu;
        double acc = 0.0;
        if (DYN == 0) {
          for (int c = 0; c < P.nu; ++c) acc = fma(M.B[threadIdx.x * MAX_NU + c], __ldg(u + c), acc);
        } else {
          const int ui = (threadIdx.x == 0 || threadIdx.x == 3) ? 0 : 1;
          acc = M.qt[4 + threadIdx.x] * __ldg(u + ui);
        }
        sh.bu[threadIdx.x] = acc;
      }
      __syncthreads();
      double bu[NX];
#pragma unroll
      for (int r = 0; r < NX; ++r) bu[r] = sh.bu[r];
      double s4[NX];
#pragma unroll
      for (int r = 0; r < NX; ++r) s4[r] = 0.0;
      for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
        double x[NX], z[NX];
        load_x<NX>(P.x, P.ld, i, x);
        noise_vector<NX, NY>(P.key, (uint32_t)t_index, i, z, sh);
        dynamics_mean<NX, NY, DYN>(M, sh, bu, t, x);
#pragma unroll
        for (int r = 0; r < NX; ++r) { x[r] += z[r]; s4[r] += x[r]; }   // f(xi,u,p,t) .+ noise_samples[i]  enkf.jl:256
        store_x<NX>(P.x, P.ld, i, x);
      }
      enkf_grid_sum<NX>(P, s4, red, tot, bar_target, seq);
      if (P.inflation > 1.0) {                       // enkf.jl:261-266
        double xb[NX];
#pragma unroll
        for (int r = 0; r < NX; ++r) { xb[r] = s4[r] * invN; s4[r] = 0.0; }
        for (int i = beg + threadIdx.x; i < end; i += BLOCK) {
          double x[NX];
          load_x<NX>(P.x, P.ld, i, x);
#pragma unroll
          for (int r = 0; r < NX; ++r) { x[r] = xb[r] + P.inflation * (x[r] - xb[r]); s4[r] += x[r]; }
          store_x<NX>(P.x, P.ld, i, x);
        }
        enkf_grid_sum<NX>(P, s4, red, tot, bar_target, seq);
      }
      t_index += 1;                                  // enkf.jl:268
      covariance_pass(s4);
    }
  }
  if (!P.do_correct && P.T > 0) put_state(P.o_x, P.o_R, 0);
  // every block has read the incoming state before block 0 overwrites it
  grid_barrier(P.bar, (unsigned)P.nblocks, bar_target);
  if (out0) {
    P.st[0] = ll_total;
    P.st[1] = (double)t_index;
#pragma unroll
    for (int r = 0; r < NX; ++r) P.st[2 + r] = mean[r];
#pragma unroll
    for (int q = 0; q < NX * NX; ++q) P.st[2 + NX + q] = cov[q];
    P.st[2 + NX + NX * NX] = (double)status;
  }
}

}  // namespace llpf
