// kalman_sc.inl -- shared-covariance mode: when one prior (Q, c, R, W) serves the whole
// batch -- always the case for a KalmanODE object -- Sigma_pred, Sigma_filt, Sigma_smooth, the
// Kalman gain, the smoother gain and the sampling Cholesky factors do not depend on theta, x0 or
// the ODE right-hand side (KalmanTV.h:290-292, 335-343, 352-354, 456-465, 511-522 never read a
// mean when producing a variance).  They are computed ONCE into tables by sc_tables_block() (one
// thread per diagonal block, or one thread for a dense prior), and the batch kernels run the MEAN
// recursions only.  This is a different amount of work per system than the general mode and is
// benchmarked on its own line, never against the general-mode roofline.
//
// Tables, per time index t = 1..N (slot t-1), per block (block size q, one block of size n when dense):
//   K  : Kalman gain.  dense: K~ (m x n row-major) so mu_f = mu_p + K~^T e;  block: k (q) so
//        mu_f = mu_p + k e_a
//   C  : smoother gain T~ (q x q row-major): mu_s = mu_f + T~^T (mu_s+ - mu_p+)   (t < N)
//   L  : packed lower Cholesky: chol(Sigma_f(N)) at t = N, chol(V~_t) at t < N     (draws)
//   var: smoothed covariance in the OUTPUT layout (N+1, n, n), written by sc_tables_block
//   Sf : filtered covariance (scratch of the table pass)

// Covariance recursion of one q-dimensional block with an mq-dimensional measurement
// (dense: q = n, mq = m; block-diagonal: mq = 1).  Writes K, C, L slots and this block's part of
// var (rows/cols off .. off+q of an n x n matrix; the rest is zeroed by the caller).
template <int q, int mq, int n>
RD_HD int sc_tables_block(const double *Q, const double *R, const double *W, bool zero_H, int N, int off,
                          double *tabK, int strideK, double *tabC, int strideC, double *tabL, int strideL,
                          double *tabSf, int strideSf, double *var) {
  constexpr int QS = Sym<q>::size;
  int fail = 0;
  double S[QS];
RD_UNROLL
  for (int i = 0; i < QS; i++) S[i] = 0.0;
  for (int t = 0; t < N; t++) {              // forward: time index t -> t+1, slot t
    double Sp[QS], mu[q], y[mq > 0 ? mq : 1];
    predict_var<q, false>(Q, R, S, Sp, nullptr);
RD_UNROLL
    for (int i = 0; i < q; i++) mu[i] = 0.0;
RD_UNROLL
    for (int i = 0; i < mq; i++) y[i] = 0.0;
    int f;
    if (mq == 1)
      f = update_block<q>(W, mu, Sp, y[0], zero_H, tabK + (long)t * strideK);
    else
      f = update_dense<q, (mq > 0 ? mq : 1)>(W, mu, Sp, y, zero_H, tabK + (long)t * strideK);
    if (f && !fail) fail = f;
RD_UNROLL
    for (int i = 0; i < QS; i++) {
      S[i] = Sp[i];
      tabSf[(long)t * strideSf + i] = Sp[i];
    }
  }
  // backward: smoothed covariance, smoother gain, sampling factors
  double Ss[QS], mus[q], x[q], zeros[q];
RD_UNROLL
  for (int i = 0; i < q; i++) zeros[i] = 0.0;
  auto put_var = [&](int t) {
RD_UNROLL
    for (int j = 0; j < q; j++)
RD_UNROLL
      for (int i = 0; i < q; i++)
        var[((long)t * n + off + j) * n + off + i] = Ss[sidx<q>(i, j)];
  };
  {
    int f = smooth_terminal<q, 1, false, true, true>(zeros, tabSf + (long)(N - 1) * strideSf, zeros, mus, Ss, x,
                                                     tabL + (long)(N - 1) * strideL);
    if (f && !fail) fail = f;
    put_var(N);
  }
  for (int t = N - 1; t >= 1; t--) {
    int f = smooth_step<q, 1, false, true, true>(Q, zeros, R, zeros, tabSf + (long)(t - 1) * strideSf, zeros, mus,
                                                 Ss, x, tabC + (long)(t - 1) * strideC,
                                                 tabL + (long)(t - 1) * strideL);
    if (f && !fail) fail = f;
    put_var(t);
  }
  return fail;
}

// Table geometry of a family (doubles per time slot).
template <class Fam>
struct ScTab {
  static constexpr int n = Fam::n, m = Fam::m, NS = Fam::NS;
  static constexpr int KW = Fam::kBlockDiag ? n : m * n;
  static constexpr int CW = Fam::kBlockDiag ? Fam::kBlocks * Fam::kBlockSize * Fam::kBlockSize : n * n;
  RD_HD static long offK(int) { return 0; }
  RD_HD static long offC(int N) { return (long)N * KW; }
  RD_HD static long offL(int N) { return offC(N) + (long)N * CW; }
  RD_HD static long offSf(int N) { return offL(N) + (long)N * NS; }
  RD_HD static long offVar(int N) { return offSf(N) + (long)N * NS; }
  RD_HD static long total(int N) { return offVar(N) + (long)(N + 1) * n * n; }
};

// Build the tables: unit u = diagonal block u (block-diagonal family; u = 0..nb-1) or the whole
// system (dense family; u = 0).  tab[offVar..] must be zero-initialised by the caller.
template <class Fam>
RD_HD int sc_tables_unit(const typename Fam::PriorT &P, bool zero_H, double *tab, int u) {
  using T = ScTab<Fam>;
  const int N = P.n_eval;
  if constexpr (Fam::kBlockDiag) {
    constexpr int p = Fam::kBlockSize, PS = Sym<p>::size;
    return sc_tables_block<p, 1, Fam::n>(P.Q[u], P.R[u], P.w[u], zero_H, N, u * p, tab + T::offK(N) + u * p, T::KW,
                                         tab + T::offC(N) + u * p * p, T::CW, tab + T::offL(N) + u * PS, T::NS,
                                         tab + T::offSf(N) + u * PS, T::NS, tab + T::offVar(N));
  } else {
    return sc_tables_block<Fam::n, Fam::m, Fam::n>(P.Q, P.R, P.W, zero_H, N, 0, tab + T::offK(N), T::KW,
                                                   tab + T::offC(N), T::CW, tab + T::offL(N), T::NS,
                                                   tab + T::offSf(N), T::NS, tab + T::offVar(N));
  }
}

// mu_pred = Q mu + c for the whole state of a family
template <class Fam>
RD_HD void sc_predict_mean(const typename Fam::PriorT &P, const double *mu, double *mup) {
  if constexpr (Fam::kBlockDiag) {
    constexpr int p = Fam::kBlockSize;
RD_UNROLL
    for (int blk = 0; blk < Fam::kBlocks; blk++) predict_mean<p>(P.Q[blk], P.c + blk * p, mu + blk * p, mup + blk * p);
  } else {
    predict_mean<Fam::n>(P.Q, P.c, mu, mup);
  }
}

// out[i] = base[i] + sum_k G[k][i] d[k], G the (block-)row-major gain of one time slot
template <class Fam>
RD_HD void sc_apply_gain(const double *G, const double *base, const double *d, double *out) {
  constexpr int n = Fam::n;
  if constexpr (Fam::kBlockDiag) {
    constexpr int p = Fam::kBlockSize;
RD_UNROLL
    for (int blk = 0; blk < Fam::kBlocks; blk++)
RD_UNROLL
      for (int i = 0; i < p; i++) {
        double s = base[blk * p + i];
RD_UNROLL
        for (int k = 0; k < p; k++) s += G[blk * p * p + k * p + i] * d[blk * p + k];
        out[blk * p + i] = s;
      }
  } else {
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double s = base[i];
RD_UNROLL
      for (int k = 0; k < n; k++) s += G[k * n + i] * d[k];
      out[i] = s;
    }
  }
}

// out[i] = base[i] + sum_{k<=i} L[i][k] z[k], L the packed (block-)lower factor of one time slot
template <class Fam>
RD_HD void sc_apply_chol(const double *L, const double *base, const double *z, double *out) {
  if constexpr (Fam::kBlockDiag) {
    constexpr int p = Fam::kBlockSize, PS = Sym<p>::size;
RD_UNROLL
    for (int blk = 0; blk < Fam::kBlocks; blk++)
RD_UNROLL
      for (int i = 0; i < p; i++) {
        double s = base[blk * p + i];
RD_UNROLL
        for (int k = 0; k <= i; k++) s += L[blk * PS + sidx<p>(i, k)] * z[blk * p + k];
        out[blk * p + i] = s;
      }
  } else {
    constexpr int n = Fam::n;
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double s = base[i];
RD_UNROLL
      for (int k = 0; k <= i; k++) s += L[sidx<n>(i, k)] * z[k];
      out[i] = s;
    }
  }
}

// Mean-only forward filter of one system (KalmanODE.pyx:299-332 without the variances); stores
// mu_filt of every time index: record r = N - t, n doubles.
template <class Fam>
RD_HD void sc_forward_body(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
  using Ode = typename Fam::Ode;
  using T = ScTab<Fam>;
  constexpr int n = Fam::n, m = Fam::m;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  const int N = P.n_eval;
  const double *tabK = a.tab + T::offK(N);
  double mu[n], th[NT];
RD_UNROLL
  for (int i = 0; i < n; i++) mu[i] = a.x0[b * n + i];
RD_UNROLL
  for (int i = 0; i < NT; i++) th[i] = (Ode::n_theta > 0) ? a.theta[b * Ode::n_theta + i] : 0.0;
  for (int t = 0; t < N; t++) {
    double mup[n], y[m];
    sc_predict_mean<Fam>(P, mu, mup);
    const double tt = P.tmin + (P.tmax - P.tmin) * (t + 1) / N;
    Ode::template eval<n>(mup, tt, th, y);
    const double *K = tabK + (long)t * T::KW;
    if constexpr (Fam::kBlockDiag) {
      constexpr int p = Fam::kBlockSize;
RD_UNROLL
      for (int blk = 0; blk < Fam::kBlocks; blk++) {
        double e = y[blk];
RD_UNROLL
        for (int j = 0; j < p; j++) e -= P.w[blk][j] * mup[blk * p + j];
RD_UNROLL
        for (int j = 0; j < p; j++) mu[blk * p + j] = mup[blk * p + j] + K[blk * p + j] * e;
      }
    } else {
      double e[m];
RD_UNROLL
      for (int q = 0; q < m; q++) {
        double s = 0.0;
RD_UNROLL
        for (int k = 0; k < n; k++) s += P.W[q * n + k] * mup[k];
        e[q] = y[q] - s;
      }
RD_UNROLL
      for (int j = 0; j < n; j++) {
        double s = mup[j];
RD_UNROLL
        for (int q = 0; q < m; q++) s += K[q * n + j] * e[q];
        mu[j] = s;
      }
    }
    double *h = a.hist + hist_offset(b, N - (t + 1), N, n);
RD_UNROLL
    for (int i = 0; i < n; i++) h[i * kTile] = mu[i];
  }
  if (a.status) a.status[b] = a.tab_fail;
}

// Mean-only backward pass: smoothed means (DO_MEAN), draws (DO_SIM), or the thinned
// log-likelihood of either (LOGLIK).  Direct per-thread global stores (the CUDA kernel stages
// them through shared memory instead; see kalman_sc.cuh).
template <class Fam, int STRIDE, bool DO_MEAN, bool DO_SIM>
RD_HD void sc_backward_terminal(const typename Fam::PriorT &P, const SolveArgs &a, const double *rec,
                                const double *zt, double *mus, double *x) {
  using T = ScTab<Fam>;
  constexpr int n = Fam::n;
  const int N = P.n_eval;
  double muf[n];
RD_UNROLL
  for (int i = 0; i < n; i++) muf[i] = rec[i * STRIDE];
  if (DO_MEAN) {
RD_UNROLL
    for (int i = 0; i < n; i++) mus[i] = muf[i];
  }
  if (DO_SIM) {
    double z[n];
RD_UNROLL
    for (int i = 0; i < n; i++) z[i] = zt[i];
    sc_apply_chol<Fam>(a.tab + T::offL(N) + (long)(N - 1) * Fam::NS, muf, z, x);
  }
}

template <class Fam, int STRIDE, bool DO_MEAN, bool DO_SIM>
RD_HD void sc_backward_step(const typename Fam::PriorT &P, const SolveArgs &a, int t, const double *rec,
                            const double *zt, double *mus, double *x) {
  using T = ScTab<Fam>;
  constexpr int n = Fam::n;
  const int N = P.n_eval;
  double muf[n], mup[n], d[n];
RD_UNROLL
  for (int i = 0; i < n; i++) muf[i] = rec[i * STRIDE];
  sc_predict_mean<Fam>(P, muf, mup);
  const double *G = a.tab + T::offC(N) + (long)(t - 1) * T::CW;
  if (DO_MEAN) {
RD_UNROLL
    for (int i = 0; i < n; i++) d[i] = mus[i] - mup[i];
    sc_apply_gain<Fam>(G, muf, d, mus);
  }
  if (DO_SIM) {
    double mt[n], z[n];
RD_UNROLL
    for (int i = 0; i < n; i++) d[i] = x[i] - mup[i];
    sc_apply_gain<Fam>(G, muf, d, mt);
RD_UNROLL
    for (int i = 0; i < n; i++) z[i] = zt[i];
    sc_apply_chol<Fam>(a.tab + T::offL(N) + (long)(t - 1) * Fam::NS, mt, z, x);
  }
}

template <class Fam, bool DO_MEAN, bool DO_SIM, bool LOGLIK>
RD_HD void sc_backward_body(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
  constexpr int n = Fam::n;
  const int N = P.n_eval;
  double mus[n], x[n];
  double lp = 0.0;
  int d = LOGLIK ? a.n_obs_times - 1 : -1;
  const double *zb = DO_SIM ? a.z + b * a.z_stride : nullptr;
  for (int t = N; t >= 1; t--) {
    const double *rec = a.hist + hist_offset(b, N - t, N, n);
    const double *zt = DO_SIM ? zb + (long)(t - 1) * n : nullptr;
    if (t == N)
      sc_backward_terminal<Fam, kTile, DO_MEAN, DO_SIM>(P, a, rec, zt, mus, x);
    else
      sc_backward_step<Fam, kTile, DO_MEAN, DO_SIM>(P, a, t, rec, zt, mus, x);
    emit_time<Fam, DO_MEAN, false, DO_SIM, LOGLIK>(a, b, t, mus, nullptr, x, lp, d);
  }
  {
    double x0v[n];
RD_UNROLL
    for (int i = 0; i < n; i++) x0v[i] = a.x0[b * n + i];
    emit_time<Fam, DO_MEAN, false, DO_SIM, LOGLIK>(a, b, 0, x0v, nullptr, x0v, lp, d);
  }
  if (LOGLIK) {
    const double cst = -0.91893853320467274178 - log(a.gamma);
    a.loglik[b] = lp + cst * (double)(a.n_obs_times * a.n_obs_comp);
  }
}
