// kalman_math.inl -- FP64 building blocks of the batched KalmanODE kernels.
//
// Included TWICE by kalman_tps.cuh, once per loop policy (the RD_UNROLL macro):
//   namespace rodeo::reg  RD_UNROLL = "#pragma unroll"    every loop fully unrolled, all
//                         matrices in registers (n_state <= 6)
//   namespace rodeo::loc  RD_UNROLL = "#pragma unroll 1"  loops rolled, matrices in local
//                         memory (any n_state; slower)
// One thread owns one system: every matrix below is a private array of the calling thread.
// Symmetric matrices (Sigma, R, Cholesky factors) are PACKED, upper triangle row-major:
// element (i, j), i <= j, at i*(2n-i-1)/2 + j.  Dense matrices are row-major.
//
// Arithmetic follows the reference's KalmanTV (rodeo/eigen/KalmanTV.h): predict :280-294,
// update :324-357, smooth_mv :447-467, smooth_sim :503-533, state_sim :619-628, with the
// probde interrogation of rodeo/eigen/KalmanTVODE.h:333-341.  Only the summation order
// and the use of symmetry differ (results agree to FP64 rounding).

// In-place lower Cholesky of a packed symmetric matrix (right-looking):
// on exit A(i, j), i >= j holds L(i, j) and invd[j] = 1 / L(j, j).
// Returns 0 or 1 + index of the first non-positive pivot.
template <int n>
RD_HD int chol_packed(double *A, double *invd) {
  int fail = 0;
RD_UNROLL
  for (int j = 0; j < n; j++) {
    double d = A[sidx<n>(j, j)];
    if (!(d > 0.0) && !fail) fail = j + 1;
    double r = inv_sqrt(d);
    invd[j] = r;
    A[sidx<n>(j, j)] = d * r;
RD_UNROLL
    for (int i = j + 1; i < n; i++) A[sidx<n>(i, j)] *= r;
RD_UNROLL
    for (int k = j + 1; k < n; k++) {
      double lkj = A[sidx<n>(k, j)];
RD_UNROLL
      for (int i = k; i < n; i++) A[sidx<n>(i, k)] -= A[sidx<n>(i, j)] * lkj;
    }
  }
  return fail;
}

// x <- (L L^T)^{-1} x for a strided vector x[i*stride], L packed with invd.
template <int n>
RD_HD void chol_solve_vec(const double *L, const double *invd, double *x, int stride) {
RD_UNROLL
  for (int i = 0; i < n; i++) {
    double s = x[i * stride];
RD_UNROLL
    for (int k = 0; k < i; k++) s -= L[sidx<n>(i, k)] * x[k * stride];
    x[i * stride] = s * invd[i];
  }
RD_UNROLL
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i * stride];
RD_UNROLL
    for (int k = i + 1; k < n; k++) s -= L[sidx<n>(k, i)] * x[k * stride];
    x[i * stride] = s * invd[i];
  }
}

// mu_pred = Q mu + c                                         (KalmanTV.h:287-288)
template <int n, int m>
RD_HD void predict_mean(const Prior<n, m> &P, const double *mu, double *mup) {
RD_UNROLL
  for (int i = 0; i < n; i++) {
    double s = P.c[i];
RD_UNROLL
    for (int k = 0; k < n; k++) s += P.Q[i * n + k] * mu[k];
    mup[i] = s;
  }
}

// Sp = Q S Q^T + R (packed)                                  (KalmanTV.h:290-292)
// If T != nullptr the full product T = Q S (row-major n x n) is also returned
// (it is the smoother's Q Sigma_f^T, KalmanTV.h:456).
template <int n, int m, bool KEEP_T>
RD_HD void predict_var(const Prior<n, m> &P, const double *S, double *Sp, double *T) {
RD_UNROLL
  for (int i = 0; i < n; i++) {
    double Ti[n];
RD_UNROLL
    for (int k = 0; k < n; k++) {
      double s = 0.0;
RD_UNROLL
      for (int l = 0; l < n; l++) s += P.Q[i * n + l] * S[sidx<n>(l, k)];
      Ti[k] = s;
      if (KEEP_T) T[i * n + k] = s;
    }
RD_UNROLL
    for (int j = i; j < n; j++) {
      double s = P.R[sidx<n>(i, j)];
RD_UNROLL
      for (int k = 0; k < n; k++) s += Ti[k] * P.Q[j * n + k];
      Sp[sidx<n>(i, j)] = s;
    }
  }
}

// Measurement update with the probde / schobert interrogation variance:
//   A = W Sp; H = A W^T (probde) or 0 (schobert); S = A W^T + H; K~ = S^{-1} A;
//   mu_f = mu_p + K~^T (y - W mu_p); Sf = Sp - K~^T A.
// (KalmanTVODE.h:336-339, KalmanTV.h:332-354).  Sp is updated in place to Sf, mup to muf.
template <int n, int m>
RD_HD int update(const Prior<n, m> &P, double *mu, double *S, const double *y, bool zero_H) {
  double A[m * n];
RD_UNROLL
  for (int a = 0; a < m; a++)
RD_UNROLL
    for (int j = 0; j < n; j++) {
      double s = 0.0;
RD_UNROLL
      for (int k = 0; k < n; k++) s += P.W[a * n + k] * S[sidx<n>(k, j)];
      A[a * n + j] = s;
    }
  double Sm[Sym<m>::size], invd[m];
RD_UNROLL
  for (int a = 0; a < m; a++)
RD_UNROLL
    for (int b = a; b < m; b++) {
      double s = 0.0;
RD_UNROLL
      for (int j = 0; j < n; j++) s += A[a * n + j] * P.W[b * n + j];
      Sm[sidx<m>(a, b)] = zero_H ? s : s + s;
    }
  int fail = chol_packed<m>(Sm, invd);
  double e[m];
RD_UNROLL
  for (int a = 0; a < m; a++) {
    double s = 0.0;
RD_UNROLL
    for (int k = 0; k < n; k++) s += P.W[a * n + k] * mu[k];
    e[a] = y[a] - s;
  }
  // K = S^{-1} A column by column; consume each column immediately
RD_UNROLL
  for (int j = 0; j < n; j++) {
    double kj[m];
RD_UNROLL
    for (int a = 0; a < m; a++) kj[a] = A[a * n + j];
    chol_solve_vec<m>(Sm, invd, kj, 1);
    double s = mu[j];
RD_UNROLL
    for (int a = 0; a < m; a++) s += kj[a] * e[a];
    mu[j] = s;
    // Sf(i, j) for i <= j: Sp(i, j) - sum_a K(a, i) A(a, j) = Sp(i,j) - sum_a A(a,i) K(a,j)
RD_UNROLL
    for (int i = 0; i <= j; i++) {
      double v = S[sidx<n>(i, j)];
RD_UNROLL
      for (int a = 0; a < m; a++) v -= A[a * n + i] * kj[a];
      S[sidx<n>(i, j)] = v;
    }
  }
  return fail;
}

// x = mu + chol_lower(V) z; V (packed) is destroyed.          (KalmanTV.h:619-628)
template <int n>
RD_HD int state_sim(double *x, const double *mu, double *V, const double *z) {
  double invd[n];
  int fail = chol_packed<n>(V, invd);
RD_UNROLL
  for (int i = 0; i < n; i++) {
    double s = mu[i];
RD_UNROLL
    for (int k = 0; k <= i; k++) s += V[sidx<n>(i, k)] * z[k];
    x[i] = s;
  }
  return fail;
}

