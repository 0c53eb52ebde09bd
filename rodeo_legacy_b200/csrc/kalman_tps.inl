// kalman_tps.inl -- thread-per-system forward / backward bodies (see kalman_tps.cuh), for the
// two problem families:
//   DenseFam<Ode, n>   general dense Q, R, W                     (history NE = n + n(n+1)/2)
//   BdFam<Ode, p>      block-diagonal Q, R with nb = n_meas blocks of size p, W row a inside
//                      block a (kalman_common.cuh)               (history NE = n + nb p(p+1)/2)
// A family provides: forward<IG>() -- the whole forward filter of one system; bw_terminal /
// bw_step -- the smoother recursion on a history record; var_at -- element (i, j) of the
// carried smoothed covariance.

template <class Ode_, int n_>
struct DenseFam {
  using Ode = Ode_;
  static constexpr int n = n_, m = Ode::n_meas, NS = Sym<n_>::size, NE = n_ + Sym<n_>::size;
  static constexpr bool kBlockDiag = false;
  using PriorT = Prior<n_, Ode_::n_meas>;

  template <int IG>
  RD_HD static void forward(const PriorT &P, const SolveArgs &a, long b) {
    constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
    double mu[n], S[NS], th[NT];
RD_UNROLL
    for (int i = 0; i < n; i++) mu[i] = a.x0[b * n + i];
RD_UNROLL
    for (int i = 0; i < NS; i++) S[i] = 0.0;
RD_UNROLL
    for (int i = 0; i < NT; i++) th[i] = (Ode::n_theta > 0) ? a.theta[b * Ode::n_theta + i] : 0.0;
    const int N = P.n_eval;
    const double *zb = (IG == IG_CHKREBTII) ? a.z + b * a.z_stride : nullptr;
    int fail = 0;
    for (int t = 0; t < N; t++) {
      double mup[n], Sp[NS], y[m];
      predict_mean<n>(P.Q, P.c, mu, mup);
      predict_var<n, false>(P.Q, P.R, S, Sp, nullptr);
      const double tt = P.tmin + (P.tmax - P.tmin) * (t + 1) / N;
      if (IG == IG_CHKREBTII) {  // x_state ~ N(mu_pred, Sigma_pred) with z[:, t]  (KalmanODE.pyx:13-49)
        double xs[n], V[NS], zt[n];
RD_UNROLL
        for (int i = 0; i < NS; i++) V[i] = Sp[i];
RD_UNROLL
        for (int i = 0; i < n; i++) zt[i] = zb[(long)t * n + i];
        int f = state_sim<n>(xs, mup, V, zt);
        if (f && !fail) fail = f;
        Ode::template eval<n>(xs, tt, th, y);
      } else {
        Ode::template eval<n>(mup, tt, th, y);
      }
      int f = update_dense<n, m>(P.W, mup, Sp, y, IG == IG_SCHOBERT);
      if (f && !fail) fail = f;
      double *h = a.hist + hist_offset(b, t, N, NE);
RD_UNROLL
      for (int i = 0; i < n; i++) {
        mu[i] = mup[i];
        h[i * kTile] = mup[i];
      }
RD_UNROLL
      for (int i = 0; i < NS; i++) {
        S[i] = Sp[i];
        h[(n + i) * kTile] = Sp[i];
      }
    }
    if (a.status) a.status[b] = fail;
  }

  template <bool DO_MEAN, bool DO_VAR, bool DO_SIM>
  RD_HD static int bw_terminal(const PriorT &, const double *rec, const double *zt, double *mus, double *Ss,
                               double *x) {
    return smooth_terminal<n, DO_MEAN, DO_VAR, DO_SIM>(rec, rec + n * kTile, zt, mus, Ss, x);
  }
  template <bool DO_MEAN, bool DO_VAR, bool DO_SIM>
  RD_HD static int bw_step(const PriorT &P, const double *rec, const double *zt, double *mus, double *Ss,
                           double *x) {
    return smooth_step<n, DO_MEAN, DO_VAR, DO_SIM>(P.Q, P.c, P.R, rec, rec + n * kTile, zt, mus, Ss, x);
  }
  RD_HD static double var_at(const double *Ss, int i, int j) { return Ss[sidx<n>(i, j)]; }
};

template <class Ode_, int p_>
struct BdFam {
  using Ode = Ode_;
  static constexpr int p = p_, nb = Ode::n_meas, n = p_ * Ode_::n_meas, m = Ode::n_meas;
  static constexpr int PS = Sym<p_>::size, NS = Ode_::n_meas * Sym<p_>::size, NE = n + NS;
  static constexpr bool kBlockDiag = true;
  using PriorT = PriorBD<p_, Ode_::n_meas>;

  template <int IG>
  RD_HD static void forward(const PriorT &P, const SolveArgs &a, long b) {
    constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
    double mu[n], S[NS], th[NT];
RD_UNROLL
    for (int i = 0; i < n; i++) mu[i] = a.x0[b * n + i];
RD_UNROLL
    for (int i = 0; i < NS; i++) S[i] = 0.0;
RD_UNROLL
    for (int i = 0; i < NT; i++) th[i] = (Ode::n_theta > 0) ? a.theta[b * Ode::n_theta + i] : 0.0;
    const int N = P.n_eval;
    const double *zb = (IG == IG_CHKREBTII) ? a.z + b * a.z_stride : nullptr;
    int fail = 0;
    for (int t = 0; t < N; t++) {
      double mup[n], Sp[NS], y[m];
RD_UNROLL
      for (int blk = 0; blk < nb; blk++) {
        predict_mean<p>(P.Q[blk], P.c + blk * p, mu + blk * p, mup + blk * p);
        predict_var<p, false>(P.Q[blk], P.R[blk], S + blk * PS, Sp + blk * PS, nullptr);
      }
      const double tt = P.tmin + (P.tmax - P.tmin) * (t + 1) / N;
      if (IG == IG_CHKREBTII) {  // chol of a block-diagonal matrix is the blocks' chol
        double xs[n];
RD_UNROLL
        for (int blk = 0; blk < nb; blk++) {
          double V[PS], zt[p];
RD_UNROLL
          for (int i = 0; i < PS; i++) V[i] = Sp[blk * PS + i];
RD_UNROLL
          for (int i = 0; i < p; i++) zt[i] = zb[(long)t * n + blk * p + i];
          int f = state_sim<p>(xs + blk * p, mup + blk * p, V, zt);
          if (f && !fail) fail = f;
        }
        Ode::template eval<n>(xs, tt, th, y);
      } else {
        Ode::template eval<n>(mup, tt, th, y);
      }
      double *h = a.hist + hist_offset(b, t, N, NE);
RD_UNROLL
      for (int blk = 0; blk < nb; blk++) {
        int f = update_block<p>(P.w[blk], mup + blk * p, Sp + blk * PS, y[blk], IG == IG_SCHOBERT);
        if (f && !fail) fail = blk + 1;
      }
RD_UNROLL
      for (int i = 0; i < n; i++) {
        mu[i] = mup[i];
        h[i * kTile] = mup[i];
      }
RD_UNROLL
      for (int i = 0; i < NS; i++) {
        S[i] = Sp[i];
        h[(n + i) * kTile] = Sp[i];
      }
    }
    if (a.status) a.status[b] = fail;
  }

  template <bool DO_MEAN, bool DO_VAR, bool DO_SIM>
  RD_HD static int bw_terminal(const PriorT &, const double *rec, const double *zt, double *mus, double *Ss,
                               double *x) {
    int fail = 0;
RD_UNROLL
    for (int blk = 0; blk < nb; blk++) {
      int f = smooth_terminal<p, DO_MEAN, DO_VAR, DO_SIM>(rec + blk * p * kTile, rec + (n + blk * PS) * kTile,
                                                          DO_SIM ? zt + blk * p : nullptr, mus + blk * p,
                                                          Ss + blk * PS, x + blk * p);
      if (f && !fail) fail = blk * p + f;
    }
    return fail;
  }
  template <bool DO_MEAN, bool DO_VAR, bool DO_SIM>
  RD_HD static int bw_step(const PriorT &P, const double *rec, const double *zt, double *mus, double *Ss,
                           double *x) {
    int fail = 0;
RD_UNROLL
    for (int blk = 0; blk < nb; blk++) {
      int f = smooth_step<p, DO_MEAN, DO_VAR, DO_SIM>(P.Q[blk], P.c + blk * p, P.R[blk], rec + blk * p * kTile,
                                                      rec + (n + blk * PS) * kTile,
                                                      DO_SIM ? zt + blk * p : nullptr, mus + blk * p,
                                                      Ss + blk * PS, x + blk * p);
      if (f && !fail) fail = blk * p + f;
    }
    return fail;
  }
  RD_HD static double var_at(const double *Ss, int i, int j) {
    return (i / p == j / p) ? Ss[(i / p) * PS + sidx<p>(i % p, j % p)] : 0.0;
  }
};

// Gaussian log-density pieces accumulated over observed components at one data time.
template <int n>
RD_HD double loglik_terms(const SolveArgs &a, int d, const double *X) {
  double lp = 0.0;
  for (int k = 0; k < a.n_obs_comp; k++) {
    int si = a.state_index[k];
    double xv = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) xv = (i == si) ? X[i] : xv;  // keeps X in registers
    double r = (a.Y[d * a.n_obs_comp + k] - xv) / a.gamma;
    lp -= 0.5 * r * r;
  }
  return lp;
}

// Generic backward pass: plain global loads and direct (per-thread) global stores.  Used by
// the rolled family, by every log-likelihood kernel and by tests/hostsim; the register
// families' solve_mv / solve_sim / solve kernels use the shared-memory staged version in
// kalman_staged.cuh.  DO_MEAN / DO_VAR: smooth_mv; DO_SIM: smooth_sim; LOGLIK: no per-step
// output, accumulate the thinned log-likelihood of the mean (DO_MEAN) or draw (DO_SIM).
template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
RD_HD void backward_body(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
  constexpr int n = Fam::n, NE = Fam::NE;
  const int N = P.n_eval;
  const long T1 = (long)N + 1;
  double mus[n], Ss[Fam::NS], x[n];
  int fail = 0;
  double lp = 0.0;
  int d = LOGLIK ? a.n_obs_times - 1 : -1;
  const double *zb = DO_SIM ? a.z + b * a.z_stride : nullptr;
  double *mu_o = (DO_MEAN && !LOGLIK) ? a.mu_out + b * T1 * n : nullptr;
  double *var_o = (DO_VAR && !LOGLIK) ? a.var_out + b * T1 * n * n : nullptr;
  double *x_o = (DO_SIM && !LOGLIK) ? a.x_out + b * T1 * n : nullptr;
  for (int t = N; t >= 1; t--) {
    const double *rec = a.hist + hist_offset(b, t - 1, N, NE);
    const double *zt = DO_SIM ? zb + (long)(t - 1) * n : nullptr;
    int f = (t == N) ? Fam::template bw_terminal<DO_MEAN, DO_VAR, DO_SIM>(P, rec, zt, mus, Ss, x)
                     : Fam::template bw_step<DO_MEAN, DO_VAR, DO_SIM>(P, rec, zt, mus, Ss, x);
    if (f && !fail) fail = f;
    if (!LOGLIK) {
      if (DO_MEAN) {
RD_UNROLL
        for (int i = 0; i < n; i++) mu_o[(long)t * n + i] = mus[i];
      }
      if (DO_SIM) {
RD_UNROLL
        for (int i = 0; i < n; i++) x_o[(long)t * n + i] = x[i];
      }
      if (DO_VAR) {
RD_UNROLL
        for (int j = 0; j < n; j++)
RD_UNROLL
          for (int i = 0; i < n; i++) var_o[((long)t * n + j) * n + i] = Fam::var_at(Ss, i, j);
      }
    } else {
      while (d >= 0 && a.thin_index[d] == t) {
        lp += loglik_terms<n>(a, d, DO_SIM ? x : mus);
        d--;
      }
    }
  }
  // time index 0: the initial state is known (KalmanODE.pyx:445, 409); Sigma_0 = 0
  if (!LOGLIK) {
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double v = a.x0[b * n + i];
      if (DO_MEAN) mu_o[i] = v;
      if (DO_SIM) x_o[i] = v;
    }
    if (DO_VAR) {
RD_UNROLL
      for (int i = 0; i < n * n; i++) var_o[i] = 0.0;
    }
  } else {
    double x0v[n];
RD_UNROLL
    for (int i = 0; i < n; i++) x0v[i] = a.x0[b * n + i];
    while (d >= 0 && a.thin_index[d] == 0) {
      lp += loglik_terms<n>(a, d, x0v);
      d--;
    }
    // constant of the normal density: -(log sqrt(2 pi) + log gamma) per observation
    const double cst = -0.91893853320467274178 - log(a.gamma);
    a.loglik[b] = lp + cst * (double)(a.n_obs_times * a.n_obs_comp);
  }
  if (a.status && fail && a.status[b] == 0) a.status[b] = fail;
}
