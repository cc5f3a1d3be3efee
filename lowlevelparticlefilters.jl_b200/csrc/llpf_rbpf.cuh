// llpf_rbpf.cuh — device side of the Rao-Blackwellized ("marginalized") particle filter, reference src/rbpf.jl.
//
// The reference's particle is RBParticle(xn, xl, R) (rbpf.jl:1-5): a nonlinear state, and mean + covariance of a Kalman
// filter for the conditionally linear state
//     xn+ = fn(xn,u,p,t) + An xl + wn,   xl+ = A xl + B u + wl,   y = g(xn,u,p,t) + C xl + e      (rbpf.jl:88-93).
// Here a particle is the vector [xn (NXN); xl (NXL); lower triangle of R, row-major (NXL(NXL+1)/2)] of the f64 engine,
// and the four functions below are what the run-time-compiled user model (llpf_create_user with LLPF_USER_STATE_HOOKS,
// include/llpf.h) calls from its dynamics / add_noise / loglik / correct_state; the host (rbpf.py, julia/LLPFB200.jl)
// generates that source from the matrices and the two device snippets fn and g.  Everything is inlined into the fused
// sweep of k_engine<NX, NY, LLPF_DYN_USER>; the loops below unroll completely (dimensions are template constants, the
// matrices are literals of the generated source).
//
// Followed literally, including what rbpf.jl does NOT do: the new nonlinear state is drawn with covariance R1n around
// fn + An xl (rbpf.jl:222-223), not with An R An' + R1n.
#pragma once

namespace llpf_rbpf {

template <int NXN, int NXL, int NY, int NU>
struct Consts {
  double A[NXL][NXL];
  double B[NXL][NU > 0 ? NU : 1];
  double C[NY][NXL];
  double An[NXN][NXL];
  double R1l[NXL][NXL];
  double R1n[NXN][NXN];
  double R2[NY][NY];
  int zeroAn, zeroC;   // iszero(An) / iszero(C): rbpf.jl:179,247
};

template <int NXN, int NXL>
__device__ __forceinline__ int tri(int r, int c) { return NXN + NXL + r * (r + 1) / 2 + c; }   // r >= c

template <int NXN, int NXL, int NX>
__device__ __forceinline__ void load_cov(const double (&x)[NX], double (&R)[NXL][NXL]) {
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) { R[r][c] = x[tri<NXN, NXL>(r, c)]; R[c][r] = R[r][c]; }
}
template <int NXN, int NXL, int NX>
__device__ __forceinline__ void store_cov(double (&x)[NX], const double (&R)[NXL][NXL]) {
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) x[tri<NXN, NXL>(r, c)] = R[r][c];
}

// lower Cholesky factor of a symmetric positive definite N x N matrix (cholesky(Symmetric(S)), filtering.jl:122)
template <int N>
__device__ __forceinline__ void chol(const double (&S)[N][N], double (&L)[N][N]) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double d = S[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    d = sqrt(d);
    L[j][j] = d;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double v = S[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v / d;
    }
#pragma unroll
    for (int i = 0; i < j; ++i) L[i][j] = 0.0;
  }
}
// X = M / cholesky(S): every row of M solved against L L'
template <int M_, int N>
__device__ __forceinline__ void right_divide(const double (&M)[M_][N], const double (&L)[N][N], double (&X)[M_][N]) {
#pragma unroll
  for (int r = 0; r < M_; ++r) {
    double yv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double v = M[r][i];
#pragma unroll
      for (int k = 0; k < i; ++k) v -= L[i][k] * yv[k];
      yv[i] = v / L[i][i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
      double v = yv[i];
#pragma unroll
      for (int k = i + 1; k < N; ++k) v -= L[k][i] * X[r][k];
      X[r][i] = v / L[i][i];
    }
  }
}
// extended_logpdf(SimpleMvNormal(PDMat(S, chol)), e) = -(k log 2pi + logdet S)/2 - |L \ e|^2 / 2   (utils.jl:252-257)
template <int N>
__device__ __forceinline__ double logpdf_chol(const double (&L)[N][N], const double (&e)[N]) {
  double v[N], q = 0.0, ld = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double acc = e[i];
#pragma unroll
    for (int k = 0; k < i; ++k) acc -= L[i][k] * v[k];
    v[i] = acc / L[i][i];
    q += v[i] * v[i];
    ld += log(L[i][i]);
  }
  return -(N * 1.8378770664093453 + 2.0 * ld) / 2 - q / 2;
}

// ---- predict!(pf::RBPF, ...) rbpf.jl:184-229, the part that needs no noise: xn <- fn, xl <- A xl + B u, R <- R1 ----------
template <int NXN, int NXL, int NY, int NU, int NX>
__device__ __forceinline__ void predict_mean(double (&x)[NX], const double (&fn)[NXN], const double* u,
                                             const Consts<NXN, NXL, NY, NU>& k) {
  double xl[NXL], R[NXL][NXL];
#pragma unroll
  for (int r = 0; r < NXL; ++r) xl[r] = x[NXN + r];
  load_cov<NXN, NXL>(x, R);
  double AR[NXL][NXL], R1[NXL][NXL];
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int c = 0; c < NXL; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < NXL; ++m) acc += k.A[r][m] * R[m][c];
      AR[r][c] = acc;
    }
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int c = 0; c < NXL; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < NXL; ++m) acc += AR[r][m] * k.A[c][m];
      R1[r][c] = acc + k.R1l[r][c];                                   // Al*R*Al' + R1l   :209 / :219
    }
  if (!k.zeroAn) {
    // Nt = An*R*An' + R1n ; L = Al*R*An' / Nt ; R1 -= L*Nt*L'        :217-219
    double RAn[NXL][NXN], Nt[NXN][NXN], ARAn[NXL][NXN], Ln[NXN][NXN], L[NXL][NXN];
#pragma unroll
    for (int r = 0; r < NXL; ++r)
#pragma unroll
      for (int c = 0; c < NXN; ++c) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int m = 0; m < NXL; ++m) { a += R[r][m] * k.An[c][m]; b += AR[r][m] * k.An[c][m]; }
        RAn[r][c] = a; ARAn[r][c] = b;
      }
#pragma unroll
    for (int r = 0; r < NXN; ++r)
#pragma unroll
      for (int c = 0; c < NXN; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < NXL; ++m) acc += k.An[r][m] * RAn[m][c];
        Nt[r][c] = acc + k.R1n[r][c];
      }
    chol<NXN>(Nt, Ln);
    right_divide<NXL, NXN>(ARAn, Ln, L);
    double LN[NXL][NXN];
#pragma unroll
    for (int r = 0; r < NXL; ++r)
#pragma unroll
      for (int c = 0; c < NXN; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < NXN; ++m) acc += L[r][m] * Nt[m][c];
        LN[r][c] = acc;
      }
#pragma unroll
    for (int r = 0; r < NXL; ++r)
#pragma unroll
      for (int c = 0; c < NXL; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < NXN; ++m) acc += LN[r][m] * L[c][m];
        R1[r][c] -= acc;
      }
  }
#pragma unroll
  for (int r = 0; r < NXN; ++r) x[r] = fn[r];
#pragma unroll
  for (int r = 0; r < NXL; ++r) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int m = 0; m < NXL; ++m) a += k.A[r][m] * xl[m];
    if (NU > 0) {
#pragma unroll
      for (int m = 0; m < NU; ++m) b += k.B[r][m] * u[m];
    }
    x[NXN + r] = a + b;                                               // Al*xl + Bl*u   :207 / :225
  }
  store_cov<NXN, NXL>(x, R1);
}

// ---- how the drawn noise enters (rbpf.jl:203 / :221-225): xn += noise, or z = An xl + noise; xn += z; xl += L (z - An xl) ----
template <int NXN, int NXL, int NY, int NU, int NX>
__device__ __forceinline__ void add_noise(double (&x)[NX], const double (&xprev)[NX], const double (&nz)[NX],
                                          const Consts<NXN, NXL, NY, NU>& k) {
  if (k.zeroAn) {
#pragma unroll
    for (int r = 0; r < NXN; ++r) x[r] += nz[r];
    return;
  }
  double xl[NXL], R[NXL][NXL];
#pragma unroll
  for (int r = 0; r < NXL; ++r) xl[r] = xprev[NXN + r];
  load_cov<NXN, NXL>(xprev, R);
  double RAn[NXL][NXN], ARAn[NXL][NXN], Nt[NXN][NXN], Ln[NXN][NXN], L[NXL][NXN];
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int c = 0; c < NXN; ++c) {
      double a = 0.0;
#pragma unroll
      for (int m = 0; m < NXL; ++m) a += R[r][m] * k.An[c][m];
      RAn[r][c] = a;
    }
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int c = 0; c < NXN; ++c) {
      double a = 0.0;
#pragma unroll
      for (int m = 0; m < NXL; ++m) a += k.A[r][m] * RAn[m][c];
      ARAn[r][c] = a;
    }
#pragma unroll
  for (int r = 0; r < NXN; ++r)
#pragma unroll
    for (int c = 0; c < NXN; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < NXL; ++m) acc += k.An[r][m] * RAn[m][c];
      Nt[r][c] = acc + k.R1n[r][c];
    }
  chol<NXN>(Nt, Ln);
  right_divide<NXL, NXN>(ARAn, Ln, L);
  double d[NXN];
#pragma unroll
  for (int r = 0; r < NXN; ++r) {
    double axl = 0.0;
#pragma unroll
    for (int m = 0; m < NXL; ++m) axl += k.An[r][m] * xl[m];
    const double z = axl + nz[r];                                     // :222
    x[r] += z;                                                        // :223
    d[r] = z - axl;
  }
#pragma unroll
  for (int r = 0; r < NXL; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < NXN; ++m) acc += L[r][m] * d[m];
    x[NXN + r] += acc;                                                // :225
  }
}

// innovation of the particle's Kalman filter: e = (y - yn) - C xl, S = symmetrize(C R C') + R2 -> chol   (filtering.jl:102-122)
template <int NXN, int NXL, int NY, int NU, int NX>
__device__ __forceinline__ void innovation(const double (&x)[NX], const double (&yn)[NY], const double* y,
                                           const Consts<NXN, NXL, NY, NU>& k, double (&e)[NY], double (&Ls)[NY][NY],
                                           double (&RCt)[NXL][NY]) {
  double R[NXL][NXL];
  load_cov<NXN, NXL>(x, R);
#pragma unroll
  for (int a = 0; a < NY; ++a) {
    double cx = 0.0;
#pragma unroll
    for (int m = 0; m < NXL; ++m) cx += k.C[a][m] * x[NXN + m];
    e[a] = (y[a] - yn[a]) - cx;
  }
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int a = 0; a < NY; ++a) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < NXL; ++m) acc += R[r][m] * k.C[a][m];
      RCt[r][a] = acc;
    }
  double S0[NY][NY], S[NY][NY];
#pragma unroll
  for (int a = 0; a < NY; ++a)
#pragma unroll
    for (int b = 0; b < NY; ++b) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < NXL; ++m) acc += k.C[a][m] * RCt[m][b];
      S0[a][b] = acc;
    }
#pragma unroll
  for (int a = 0; a < NY; ++a)
#pragma unroll
    for (int b = 0; b < NY; ++b) S[a][b] = 0.5 * (S0[a][b] + S0[b][a]) + k.R2[a][b];
  chol<NY>(S, Ls);
}

// w[i] += ll of correct!(kf, u, y - yn)  (rbpf.jl:263-271), or logpdf(R2, y - yh) when C == 0 (:274-275)
template <int NXN, int NXL, int NY, int NU, int NX>
__device__ __forceinline__ double loglik(const double (&x)[NX], const double (&yn)[NY], const double* y,
                                         const Consts<NXN, NXL, NY, NU>& k) {
  double e[NY], Ls[NY][NY];
  if (k.zeroC) {
#pragma unroll
    for (int a = 0; a < NY; ++a) e[a] = y[a] - yn[a];
    chol<NY>(k.R2, Ls);
    return logpdf_chol<NY>(Ls, e);
  }
  double RCt[NXL][NY];
  innovation<NXN, NXL, NY, NU, NX>(x, yn, y, k, e, Ls, RCt);
  return logpdf_chol<NY>(Ls, e);
}

// the particle mutation of correct!: K = R C' / chol(S); xl += K e; R = symmetrize((I - K C) R)   (filtering.jl:124-126)
template <int NXN, int NXL, int NY, int NU, int NX>
__device__ __forceinline__ void correct_state(double (&x)[NX], const double (&yn)[NY], const double* y,
                                              const Consts<NXN, NXL, NY, NU>& k) {
  if (k.zeroC) return;
  double e[NY], Ls[NY][NY], RCt[NXL][NY], K[NXL][NY], R[NXL][NXL], R2[NXL][NXL];
  innovation<NXN, NXL, NY, NU, NX>(x, yn, y, k, e, Ls, RCt);
  right_divide<NXL, NY>(RCt, Ls, K);
  load_cov<NXN, NXL>(x, R);
#pragma unroll
  for (int r = 0; r < NXL; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < NY; ++a) acc += K[r][a] * e[a];
    x[NXN + r] += acc;
  }
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int c = 0; c < NXL; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < NXL; ++m) {
        double kc = 0.0;
#pragma unroll
        for (int a = 0; a < NY; ++a) kc += K[r][a] * k.C[a][m];
        acc += ((r == m ? 1.0 : 0.0) - kc) * R[m][c];
      }
      R2[r][c] = acc;
    }
  double Rs[NXL][NXL];
#pragma unroll
  for (int r = 0; r < NXL; ++r)
#pragma unroll
    for (int c = 0; c < NXL; ++c) Rs[r][c] = 0.5 * (R2[r][c] + R2[c][r]);
  store_cov<NXN, NXL>(x, Rs);
}

}  // namespace llpf_rbpf
