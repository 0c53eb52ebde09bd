// kalman_math.inl -- FP64 building blocks of the batched KalmanODE kernels.
//
// Included TWICE by kalman_tps.cuh, once per loop policy (the RD_UNROLL macro):
//   namespace rodeo::reg  RD_UNROLL = "#pragma unroll"    every loop fully unrolled, all
//                         matrices in registers
//   namespace rodeo::loc  RD_UNROLL = "#pragma unroll 1"  loops rolled, matrices in local
//                         memory (any n_state; slow -- the dense fallback for n_state > 6)
// One thread owns one system: every matrix below is a private array of the calling thread.
// Symmetric matrices (Sigma, R, Cholesky factors) are PACKED, upper triangle row-major:
// element (i, j), i <= j, at i*(2n-i-1)/2 + j.  Dense matrices are row-major.  Problem
// constants (Q, c, R, W) are passed as pointers INTO the __grid_constant__ kernel parameter,
// so after unrolling each coefficient is an immediate constant-bank operand.
//
// Arithmetic follows the reference's KalmanTV (rodeo/eigen/KalmanTV.h): predict :280-294,
// update :324-357, smooth_mv :447-467, smooth_sim :503-533, state_sim :619-628, with the
// probde interrogation of rodeo/eigen/KalmanTVODE.h:333-341.  Only the summation order,
// the use of symmetry / block structure and the association of two products differ (results
// agree to FP64 rounding): the mean-only step solves Sigma_p v = mu_s+ - mu_p+ once instead of
// forming the gain, and the draw goes through U = L^{-1} Q Sigma_f (see smooth_step).

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
RD_UNROLL_IN
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
RD_UNROLL_IN
    for (int k = 0; k < i; k++) s -= L[sidx<n>(i, k)] * x[k * stride];
    x[i * stride] = s * invd[i];
  }
RD_UNROLL
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i * stride];
RD_UNROLL_IN
    for (int k = i + 1; k < n; k++) s -= L[sidx<n>(k, i)] * x[k * stride];
    x[i * stride] = s * invd[i];
  }
}

// The two halves of chol_solve_vec: x <- L^{-1} x and x <- L^{-T} x.
template <int n>
RD_HD void chol_fwd_vec(const double *L, const double *invd, double *x, int stride) {
RD_UNROLL
  for (int i = 0; i < n; i++) {
    double s = x[i * stride];
RD_UNROLL_IN
    for (int k = 0; k < i; k++) s -= L[sidx<n>(i, k)] * x[k * stride];
    x[i * stride] = s * invd[i];
  }
}
template <int n>
RD_HD void chol_bwd_vec(const double *L, const double *invd, double *x, int stride) {
RD_UNROLL
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i * stride];
RD_UNROLL_IN
    for (int k = i + 1; k < n; k++) s -= L[sidx<n>(k, i)] * x[k * stride];
    x[i * stride] = s * invd[i];
  }
}

// Structured priors (detected on the host, api.cu: is_structured).  UT: Q is upper triangular (what ibm_init
// produces: a Toeplitz upper-triangular transition block); the terms with an exactly-zero coefficient are
// skipped, the others are accumulated in the same order, so for finite data the results are bit-identical
// to the general code (x * 0 = 0 and s + 0 = s exactly).  WK >= 0: w = e_WK (zero_pad puts a single 1 per
// row of wgt_meas, at the ODE's order): products with w reduce to picking element WK.

// mu_pred = Q mu + c                                         (KalmanTV.h:287-288)
template <int n, bool UT = false>
RD_HD void predict_mean(const double *Q, const double *c, const double *mu, double *mup) {
RD_UNROLL
  for (int i = 0; i < n; i++) {
    double s = c[i];
RD_UNROLL_IN
    for (int k = UT ? i : 0; k < n; k++) s += Q[i * n + k] * mu[k];
    mup[i] = s;
  }
}

// Sp = Q S Q^T + R (packed)                                  (KalmanTV.h:290-292)
// With KEEP_T the full product T = Q S (row-major n x n) is also returned (it is the
// smoother's Q Sigma_f^T, KalmanTV.h:456).
template <int n, bool KEEP_T, bool UT = false>
RD_HD void predict_var(const double *Q, const double *R, const double *S, double *Sp, double *T) {
RD_UNROLL
  for (int i = 0; i < n; i++) {
    double Ti[n];
RD_UNROLL
    for (int k = 0; k < n; k++) {
      double s = 0.0;
RD_UNROLL_IN
      for (int l = UT ? i : 0; l < n; l++) s += Q[i * n + l] * S[sidx<n>(l, k)];
      Ti[k] = s;
      if (KEEP_T) T[i * n + k] = s;
    }
RD_UNROLL
    for (int j = i; j < n; j++) {
      double s = R[sidx<n>(i, j)];
RD_UNROLL_IN
      for (int k = UT ? j : 0; k < n; k++) s += Ti[k] * Q[j * n + k];
      Sp[sidx<n>(i, j)] = s;
    }
  }
}

// Dense measurement update with the probde / schobert interrogation variance:
//   A = W Sp; H = A W^T (probde) or 0 (schobert); S = A W^T + H; K~ = S^{-1} A;
//   mu_f = mu_p + K~^T (y - W mu_p); Sf = Sp - K~^T A.
// (KalmanTVODE.h:336-339, KalmanTV.h:332-354).  S (Sp) and mu (mup) are updated in place.
template <int n, int m>
RD_HD int update_dense(const double *W, double *mu, double *S, const double *y, bool zero_H,
                       double *dumpK = nullptr) {
  double A[m * n];
RD_UNROLL
  for (int a = 0; a < m; a++)
RD_UNROLL
    for (int j = 0; j < n; j++) {
      double s = 0.0;
RD_UNROLL_IN
      for (int k = 0; k < n; k++) s += W[a * n + k] * S[sidx<n>(k, j)];
      A[a * n + j] = s;
    }
  double Sm[Sym<m>::size], invd[m];
RD_UNROLL
  for (int a = 0; a < m; a++)
RD_UNROLL
    for (int b = a; b < m; b++) {
      double s = 0.0;
RD_UNROLL_IN
      for (int j = 0; j < n; j++) s += A[a * n + j] * W[b * n + j];
      Sm[sidx<m>(a, b)] = zero_H ? s : s + s;
    }
  int fail = chol_packed<m>(Sm, invd);
  double e[m];
RD_UNROLL
  for (int a = 0; a < m; a++) {
    double s = 0.0;
RD_UNROLL_IN
    for (int k = 0; k < n; k++) s += W[a * n + k] * mu[k];
    e[a] = y[a] - s;
  }
  // K = S^{-1} A column by column; consume each column immediately
RD_UNROLL
  for (int j = 0; j < n; j++) {
    double kj[m];
RD_UNROLL
    for (int a = 0; a < m; a++) kj[a] = A[a * n + j];
    chol_solve_vec<m>(Sm, invd, kj, 1);
    if (dumpK) {                       // shared-covariance tables: K~ (m x n, row-major)
RD_UNROLL
      for (int a = 0; a < m; a++) dumpK[a * n + j] = kj[a];
    }
    double s = mu[j];
RD_UNROLL_IN
    for (int a = 0; a < m; a++) s += kj[a] * e[a];
    mu[j] = s;
    // Sf(i, j), i <= j: Sp(i, j) - sum_a K(a, i) A(a, j) = Sp(i, j) - sum_a A(a, i) K(a, j)
RD_UNROLL
    for (int i = 0; i <= j; i++) {
      double v = S[sidx<n>(i, j)];
RD_UNROLL_IN
      for (int a = 0; a < m; a++) v -= A[a * n + i] * kj[a];
      S[sidx<n>(i, j)] = v;
    }
  }
  return fail;
}

// The same update for ONE diagonal block with a scalar measurement  y_a = w^T x_a:
//   A = w^T Sp; s = A w (+ A w under probde); K = A / s; mu_f = mu_p + K (y_a - w^T mu_p);
//   Sf = Sp - A^T K.   The m x m Cholesky of the dense update degenerates to one division.
// The two halves are separate routines so that a kernel can start the division (which needs only Sp) before the
// right-hand side is known (kalman_wpb.cuh); update_block is their composition, so every kernel runs the same
// operations in the same order.
template <int p, int WK = -1>
RD_HD int update_block_gain(const double *w, const double *S, bool zero_H, double *A, double *inv) {
  double s = 0.0;
  if (WK >= 0) {                 // w = e_WK: A = row WK of Sp, s = Sp(WK, WK)
    constexpr int K = WK >= 0 ? WK : 0;
RD_UNROLL
    for (int j = 0; j < p; j++) A[j] = S[sidx<p>(K, j)];
    s = A[K];
  } else {
RD_UNROLL
    for (int j = 0; j < p; j++) {
      double v = 0.0;
RD_UNROLL_IN
      for (int k = 0; k < p; k++) v += w[k] * S[sidx<p>(k, j)];
      A[j] = v;
    }
RD_UNROLL
    for (int j = 0; j < p; j++) s += A[j] * w[j];
  }
  if (!zero_H) s = s + s;
  *inv = 1.0 / s;
  return (s > 0.0) ? 0 : 1;
}
template <int p, int WK = -1>
RD_HD void update_block_apply(const double *w, double *mu, double *S, double ya, const double *A, double inv,
                              double *dumpK = nullptr) {
  double e = ya;
  if (WK >= 0) {                 // e = y - mu_p[WK]
    constexpr int K = WK >= 0 ? WK : 0;
    e = sub_rn(ya, mu[K]);
  } else {
RD_UNROLL
    for (int j = 0; j < p; j++) e -= w[j] * mu[j];
  }
RD_UNROLL
  for (int j = 0; j < p; j++) {
    const double kj = A[j] * inv;
    if (dumpK) dumpK[j] = kj;
    mu[j] += kj * e;
RD_UNROLL_IN
    for (int i = 0; i <= j; i++) S[sidx<p>(i, j)] -= A[i] * kj;
  }
}
template <int p, int WK = -1>
RD_HD int update_block(const double *w, double *mu, double *S, double ya, bool zero_H,
                       double *dumpK = nullptr) {
  double A[p], inv;
  const int fail = update_block_gain<p, WK>(w, S, zero_H, A, &inv);
  update_block_apply<p, WK>(w, mu, S, ya, A, inv, dumpK);
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
RD_UNROLL_IN
    for (int k = 0; k <= i; k++) s += V[sidx<n>(i, k)] * z[k];
    x[i] = s;
  }
  return fail;
}

// ---- smoother, one n-dimensional (block of a) system --------------------------------------
// History elements are read through (rec_mu, rec_S): element i at ptr[i * STRIDE] (STRIDE =
// kTile for a record in the tiled history / shared-memory ring, 1 for a thread-private array).

// Terminal time index N (KalmanODE.pyx:445-449, 410-414); zt = z[:, N-1] (this block's rows).
template <int n, int STRIDE, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
RD_HD int smooth_terminal(const double *rec_mu, const double *rec_S, const double *zt_ptr, double *mus,
                          double *Ss, double *x, double *dumpL = nullptr) {
  constexpr int NS = Sym<n>::size;
  double muf[n], Sf[NS];
  int fail = 0;
RD_UNROLL
  for (int i = 0; i < n; i++) muf[i] = rec_mu[i * STRIDE];
RD_UNROLL
  for (int i = 0; i < NS; i++) Sf[i] = rec_S[i * STRIDE];
  if (DO_MEAN) {
RD_UNROLL
    for (int i = 0; i < n; i++) mus[i] = muf[i];
  }
  if (DO_VAR) {
RD_UNROLL
    for (int i = 0; i < NS; i++) Ss[i] = Sf[i];
  }
  if (DO_SIM) {
    double zt[n];
RD_UNROLL
    for (int i = 0; i < n; i++) zt[i] = zt_ptr[i];
    fail = state_sim<n>(x, muf, Sf, zt);  // destroys Sf (not needed again): now chol_lower
    if (dumpL) {
RD_UNROLL
      for (int i = 0; i < NS; i++) dumpL[i] = Sf[i];
    }
  }
  return fail;
}

// One smoother step t+1 -> t (KalmanODE.pyx:452-461 / 417-425 / 381-392); zt_ptr = z[:, t-1].
// The filtered covariance is re-read from the record where it is needed a second time instead
// of being held in registers across the Cholesky / triangular solves (REREAD: the record lives in
// shared or global memory; with REREAD = false it is a thread-private register array and stays live).
template <bool REREAD>
struct RecReader {
  const volatile double *q;
  RD_HD explicit RecReader(const double *p) : q(p) {}
  RD_HD double operator[](int i) const { return q[i]; }
};
template <>
struct RecReader<false> {
  const double *q;
  RD_HD explicit RecReader(const double *p) : q(p) {}
  RD_HD double operator[](int i) const { return q[i]; }
};

template <int n, int STRIDE, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool REREAD = true, bool UT = false>
RD_HD int smooth_step(const double *Q, const double *c, const double *R, const double *rec_mu,
                      const double *rec_S, const double *zt_ptr, double *mus, double *Ss, double *x,
                      double *dumpT = nullptr, double *dumpL = nullptr) {
  constexpr int NS = Sym<n>::size;
  const RecReader<REREAD> rec_S2(rec_S);  // second reads: do not keep Sigma_f live
  double muf[n], mup[n], Sp[NS], T[n * n], invd[n];
  {
    double Sf[NS];
RD_UNROLL
    for (int i = 0; i < n; i++) muf[i] = rec_mu[i * STRIDE];
RD_UNROLL
    for (int i = 0; i < NS; i++) Sf[i] = rec_S[i * STRIDE];
    // predicted moments at t+1, recomputed exactly as the forward pass computed them
    predict_mean<n, UT>(Q, c, muf, mup);
    predict_var<n, true, UT>(Q, R, Sf, Sp, T);  // T = Q Sigma_f  (KalmanTV.h:456)
  }
  if (DO_VAR) {                              // D = Sigma_s(t+1) - Sigma_p(t+1)  (:460)
RD_UNROLL
    for (int i = 0; i < NS; i++) Ss[i] -= Sp[i];
  }
  int fail = chol_packed<n>(Sp, invd);       // LLT(Sigma_p)  (:457)
  if (DO_MEAN && !DO_VAR && !DO_SIM && !dumpT) {
    // Mean only (the log-likelihood smoothers): mu_s = mu_f + T^T Sigma_p^{-1} (mu_s+ - mu_p+) needs ONE solve
    // with the vector, not the n solves that form T~ = Sigma_p^{-1} T (:458-459) -- the same product associated the
    // other way (FN log-likelihood smoother: ~20 % fewer FP64 instructions; the kernel is FP64-bound).
    double v[n];
RD_UNROLL
    for (int k = 0; k < n; k++) v[k] = mus[k] - mup[k];
    chol_solve_vec<n>(Sp, invd, v, 1);
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double s = muf[i];
RD_UNROLL_IN
      for (int k = 0; k < n; k++) s += T[k * n + i] * v[k];
      mus[i] = s;
    }
    return fail;
  }
  // With Sigma_p = L L^T the gain is T~ = L^{-T} U, U = L^{-1} T (T = Q Sigma_f).  The draw needs only U:
  //   m~ = mu_f + T~^T (x+ - mu_p+) = mu_f + U^T (L^{-1} (x+ - mu_p+)),   V~ = Sigma_f - T~^T T = Sigma_f - U^T U
  // (KalmanTV.h:511-523 with the products associated the other way): no backward substitution and no second
  // Q Sigma_f for the draw-only smoothers, and V~ is formed as a symmetric difference.  Mean and variance then
  // take T~ = L^{-T} U as before (U is overwritten in place, after the draw).
RD_UNROLL
  for (int j = 0; j < n; j++) chol_fwd_vec<n>(Sp, invd, T + j, n);   // U
  if (DO_SIM) {
    double w[n], mt[n], V[NS], zt[n];
RD_UNROLL
    for (int k = 0; k < n; k++) w[k] = x[k] - mup[k];
    chol_fwd_vec<n>(Sp, invd, w, 1);
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double s = muf[i];
RD_UNROLL_IN
      for (int k = 0; k < n; k++) s += T[k * n + i] * w[k];
      mt[i] = s;
    }
RD_UNROLL
    for (int i = 0; i < NS; i++) V[i] = rec_S2[i * STRIDE];
RD_UNROLL
    for (int k = 0; k < n; k++)
RD_UNROLL
      for (int i = 0; i < n; i++)
RD_UNROLL_IN
        for (int j = i; j < n; j++) V[sidx<n>(i, j)] -= T[k * n + i] * T[k * n + j];
RD_UNROLL
    for (int i = 0; i < n; i++) zt[i] = zt_ptr[i];
    int f2 = state_sim<n>(x, mt, V, zt);
    if (f2 && !fail) fail = f2;
    if (dumpL) {                             // V now holds chol_lower(V~)
RD_UNROLL
      for (int i = 0; i < NS; i++) dumpL[i] = V[i];
    }
  }
  if (!(DO_MEAN || DO_VAR || dumpT)) return fail;
RD_UNROLL
  for (int j = 0; j < n; j++) chol_bwd_vec<n>(Sp, invd, T + j, n);   // T~ (:458)
  if (dumpT) {                               // shared-covariance tables: the smoother gain
RD_UNROLL
    for (int i = 0; i < n * n; i++) dumpT[i] = T[i];
  }
  if (DO_MEAN) {                             // (:459, 463-464)
    double dmu[n];
RD_UNROLL
    for (int k = 0; k < n; k++) dmu[k] = mus[k] - mup[k];
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double s = muf[i];
RD_UNROLL_IN
      for (int k = 0; k < n; k++) s += T[k * n + i] * dmu[k];
      mus[i] = s;
    }
  }
  if (DO_VAR) {                              // Sigma_s = Sigma_f + T~^T D T~  (:461-465)
    double Sn[NS];
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double G[n];
RD_UNROLL
      for (int l = 0; l < n; l++) {
        double s = 0.0;
RD_UNROLL_IN
        for (int k = 0; k < n; k++) s += T[k * n + i] * Ss[sidx<n>(k, l)];
        G[l] = s;
      }
RD_UNROLL
      for (int j = i; j < n; j++) {
        double s = rec_S2[sidx<n>(i, j) * STRIDE];
RD_UNROLL_IN
        for (int l = 0; l < n; l++) s += G[l] * T[l * n + j];
        Sn[sidx<n>(i, j)] = s;
      }
    }
RD_UNROLL
    for (int i = 0; i < NS; i++) Ss[i] = Sn[i];
  }
  return fail;
}
