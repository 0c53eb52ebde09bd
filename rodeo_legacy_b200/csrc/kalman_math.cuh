// kalman_math.cuh -- FP64 building blocks of the batched KalmanODE kernels.
//
// One thread owns one system ("thread-per-system"): every matrix below is a private
// array of the calling thread.  With UNROLL=true all loops are fully unrolled and the
// arrays live in registers (used for n_state <= 6); with UNROLL=false the same code
// keeps its loops rolled and the arrays in local memory (any n_state; slower -- the
// wide-state kernels in kalman_wide.cuh replace it for n_state > 6).
//
// Symmetric matrices (Sigma, R, Cholesky factors) are stored PACKED, upper triangle
// row-major: element (i, j), i <= j, at i*(2n-i-1)/2 + j.  Dense matrices are row-major.
//
// Arithmetic follows the reference's KalmanTV (rodeo/eigen/KalmanTV.h): predict
// :280-294, update :324-357, smooth_mv :447-467, smooth_sim :503-533, state_sim
// :619-628, with the probde interrogation of rodeo/eigen/KalmanTVODE.h:333-341.  Only
// the summation order and the use of symmetry differ (results agree to FP64 rounding).
#pragma once

#if defined(__CUDACC__)
#define RD_HD __host__ __device__ __forceinline__
#else
#define RD_HD inline
#endif

#include <math.h>

namespace rodeo {

template <int n>
RD_HD constexpr int sidx(int i, int j) {
  return i <= j ? i * (2 * n - i - 1) / 2 + j : j * (2 * n - j - 1) / 2 + i;
}
template <int n>
struct Sym {
  static constexpr int size = n * (n + 1) / 2;
};

// Problem constants, passed BY VALUE as a __grid_constant__ kernel parameter so that
// every coefficient is an immediate constant-bank operand of the DFMA that uses it.
template <int n, int m>
struct Prior {
  double Q[n * n];            // wgt_state, row-major
  double c[n];                // mu_state
  double R[Sym<n>::size];     // var_state, packed
  double W[m * n];            // wgt_meas, row-major
  double tmin, tmax;
  int n_eval;
};

RD_HD double inv_sqrt(double d) {
#if defined(__CUDA_ARCH__)
  return rsqrt(d);
#else
  return 1.0 / sqrt(d);
#endif
}

#define RD_U constexpr int U_ = UNROLL ? 64 : 1

// In-place lower Cholesky of a packed symmetric matrix (right-looking):
// on exit A(i, j), i >= j holds L(i, j) and invd[j] = 1 / L(j, j).
// Returns 0 or 1 + index of the first non-positive pivot.
template <int n, bool UNROLL>
RD_HD int chol_packed(double *A, double *invd) {
  RD_U;
  int fail = 0;
#pragma unroll U_
  for (int j = 0; j < n; j++) {
    double d = A[sidx<n>(j, j)];
    if (!(d > 0.0) && !fail) fail = j + 1;
    double r = inv_sqrt(d);
    invd[j] = r;
    A[sidx<n>(j, j)] = d * r;
#pragma unroll U_
    for (int i = j + 1; i < n; i++) A[sidx<n>(i, j)] *= r;
#pragma unroll U_
    for (int k = j + 1; k < n; k++) {
      double lkj = A[sidx<n>(k, j)];
#pragma unroll U_
      for (int i = k; i < n; i++) A[sidx<n>(i, k)] -= A[sidx<n>(i, j)] * lkj;
    }
  }
  return fail;
}

// x <- (L L^T)^{-1} x for a strided vector x[i*stride], L packed with invd.
template <int n, bool UNROLL>
RD_HD void chol_solve_vec(const double *L, const double *invd, double *x, int stride) {
  RD_U;
#pragma unroll U_
  for (int i = 0; i < n; i++) {
    double s = x[i * stride];
#pragma unroll U_
    for (int k = 0; k < i; k++) s -= L[sidx<n>(i, k)] * x[k * stride];
    x[i * stride] = s * invd[i];
  }
#pragma unroll U_
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i * stride];
#pragma unroll U_
    for (int k = i + 1; k < n; k++) s -= L[sidx<n>(k, i)] * x[k * stride];
    x[i * stride] = s * invd[i];
  }
}

// mu_pred = Q mu + c                                         (KalmanTV.h:287-288)
template <int n, int m, bool UNROLL>
RD_HD void predict_mean(const Prior<n, m> &P, const double *mu, double *mup) {
  RD_U;
#pragma unroll U_
  for (int i = 0; i < n; i++) {
    double s = P.c[i];
#pragma unroll U_
    for (int k = 0; k < n; k++) s += P.Q[i * n + k] * mu[k];
    mup[i] = s;
  }
}

// Sp = Q S Q^T + R (packed)                                  (KalmanTV.h:290-292)
// If T != nullptr the full product T = Q S (row-major n x n) is also returned
// (it is the smoother's Q Sigma_f^T, KalmanTV.h:456).
template <int n, int m, bool UNROLL, bool KEEP_T>
RD_HD void predict_var(const Prior<n, m> &P, const double *S, double *Sp, double *T) {
  RD_U;
#pragma unroll U_
  for (int i = 0; i < n; i++) {
    double Ti[n];
#pragma unroll U_
    for (int k = 0; k < n; k++) {
      double s = 0.0;
#pragma unroll U_
      for (int l = 0; l < n; l++) s += P.Q[i * n + l] * S[sidx<n>(l, k)];
      Ti[k] = s;
      if (KEEP_T) T[i * n + k] = s;
    }
#pragma unroll U_
    for (int j = i; j < n; j++) {
      double s = P.R[sidx<n>(i, j)];
#pragma unroll U_
      for (int k = 0; k < n; k++) s += Ti[k] * P.Q[j * n + k];
      Sp[sidx<n>(i, j)] = s;
    }
  }
}

// Measurement update with the probde / schobert interrogation variance:
//   A = W Sp; H = A W^T (probde) or 0 (schobert); S = A W^T + H; K~ = S^{-1} A;
//   mu_f = mu_p + K~^T (y - W mu_p); Sf = Sp - K~^T A.
// (KalmanTVODE.h:336-339, KalmanTV.h:332-354).  Sp is updated in place to Sf, mup to muf.
template <int n, int m, bool UNROLL>
RD_HD int update(const Prior<n, m> &P, double *mu, double *S, const double *y, bool zero_H) {
  RD_U;
  double A[m * n];
#pragma unroll U_
  for (int a = 0; a < m; a++)
#pragma unroll U_
    for (int j = 0; j < n; j++) {
      double s = 0.0;
#pragma unroll U_
      for (int k = 0; k < n; k++) s += P.W[a * n + k] * S[sidx<n>(k, j)];
      A[a * n + j] = s;
    }
  double Sm[Sym<m>::size], invd[m];
#pragma unroll U_
  for (int a = 0; a < m; a++)
#pragma unroll U_
    for (int b = a; b < m; b++) {
      double s = 0.0;
#pragma unroll U_
      for (int j = 0; j < n; j++) s += A[a * n + j] * P.W[b * n + j];
      Sm[sidx<m>(a, b)] = zero_H ? s : s + s;
    }
  int fail = chol_packed<m, UNROLL>(Sm, invd);
  double e[m];
#pragma unroll U_
  for (int a = 0; a < m; a++) {
    double s = 0.0;
#pragma unroll U_
    for (int k = 0; k < n; k++) s += P.W[a * n + k] * mu[k];
    e[a] = y[a] - s;
  }
  // K = S^{-1} A column by column; consume each column immediately
#pragma unroll U_
  for (int j = 0; j < n; j++) {
    double kj[m];
#pragma unroll U_
    for (int a = 0; a < m; a++) kj[a] = A[a * n + j];
    chol_solve_vec<m, UNROLL>(Sm, invd, kj, 1);
    double s = mu[j];
#pragma unroll U_
    for (int a = 0; a < m; a++) s += kj[a] * e[a];
    mu[j] = s;
    // Sf(i, j) for i <= j: Sp(i, j) - sum_a K(a, i) A(a, j) = Sp(i,j) - sum_a A(a,i) K(a,j)
#pragma unroll U_
    for (int i = 0; i <= j; i++) {
      double v = S[sidx<n>(i, j)];
#pragma unroll U_
      for (int a = 0; a < m; a++) v -= A[a * n + i] * kj[a];
      S[sidx<n>(i, j)] = v;
    }
  }
  return fail;
}

// x = mu + chol_lower(V) z; V (packed) is destroyed.          (KalmanTV.h:619-628)
template <int n, bool UNROLL>
RD_HD int state_sim(double *x, const double *mu, double *V, const double *z) {
  RD_U;
  double invd[n];
  int fail = chol_packed<n, UNROLL>(V, invd);
#pragma unroll U_
  for (int i = 0; i < n; i++) {
    double s = mu[i];
#pragma unroll U_
    for (int k = 0; k <= i; k++) s += V[sidx<n>(i, k)] * z[k];
    x[i] = s;
  }
  return fail;
}

}  // namespace rodeo
