// kalman_tps.inl -- thread-per-system forward / backward bodies (see kalman_tps.cuh).
template <class Ode, int n, int IG>
RD_HD void forward_body(const Prior<n, Ode::n_meas> &P, const SolveArgs &a, long b) {
  constexpr int m = Ode::n_meas, NS = Sym<n>::size, NE = n + NS;
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
    predict_mean<n, m>(P, mu, mup);
    predict_var<n, m, false>(P, S, Sp, nullptr);
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
    int f = update<n, m>(P, mup, Sp, y, IG == IG_SCHOBERT);
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

// ---- backward pass ---------------------------------------------------------------------
// State carried from time index t+1 to t.
template <int n>
struct SmoothState {
  double mus[n];            // smoothed mean            (DO_MEAN)
  double Ss[Sym<n>::size];  // smoothed covariance      (DO_VAR)
  double x[n];              // posterior draw           (DO_SIM)
  int fail;
};

// Terminal time index N (KalmanODE.pyx:445-449, 410-414).  `rec` points at this system's
// history record for time index N (element e at rec[e * kTile]); zt = z[:, N-1].
template <int n, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
RD_HD void backward_terminal(const double *rec, const double *zt_ptr, SmoothState<n> &st) {
  constexpr int NS = Sym<n>::size;
  double muf[n], Sf[NS];
RD_UNROLL
  for (int i = 0; i < n; i++) muf[i] = rec[i * kTile];
RD_UNROLL
  for (int i = 0; i < NS; i++) Sf[i] = rec[(n + i) * kTile];
  if (DO_MEAN) {
RD_UNROLL
    for (int i = 0; i < n; i++) st.mus[i] = muf[i];
  }
  if (DO_VAR) {
RD_UNROLL
    for (int i = 0; i < NS; i++) st.Ss[i] = Sf[i];
  }
  if (DO_SIM) {
    double zt[n];
RD_UNROLL
    for (int i = 0; i < n; i++) zt[i] = zt_ptr[i];
    int f = state_sim<n>(st.x, muf, Sf, zt);  // destroys Sf (not needed again)
    if (f && !st.fail) st.fail = f;
  }
}

// One smoother step t+1 -> t (KalmanODE.pyx:452-461 / 417-425 / 381-392).  `rec` is the
// history record of time index t; zt_ptr = z[:, t-1].  The filtered covariance is re-read
// from `rec` where it is needed a second time instead of being held in registers.
template <int n, int m, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
RD_HD void backward_step(const Prior<n, m> &P, const double *rec, const double *zt_ptr,
                         SmoothState<n> &st) {
  constexpr int NS = Sym<n>::size;
  const volatile double *rec2 = rec;  // second reads: do not keep Sigma_f live across the step
  double muf[n], mup[n], Sp[NS], T[n * n], invd[n];
  {
    double Sf[NS];
RD_UNROLL
    for (int i = 0; i < n; i++) muf[i] = rec[i * kTile];
RD_UNROLL
    for (int i = 0; i < NS; i++) Sf[i] = rec[(n + i) * kTile];
    // predicted moments at t+1, recomputed exactly as the forward pass computed them
    predict_mean<n, m>(P, muf, mup);
    predict_var<n, m, true>(P, Sf, Sp, T);  // T = Q Sigma_f  (KalmanTV.h:456)
  }
  if (DO_VAR) {                              // D = Sigma_s(t+1) - Sigma_p(t+1)  (:460)
RD_UNROLL
    for (int i = 0; i < NS; i++) st.Ss[i] -= Sp[i];
  }
  int f = chol_packed<n>(Sp, invd);          // LLT(Sigma_p)  (:457)
  if (f && !st.fail) st.fail = f;
RD_UNROLL
  for (int j = 0; j < n; j++) chol_solve_vec<n>(Sp, invd, T + j, n);  // T~ (:458)
  if (DO_MEAN) {                             // (:459, 463-464)
    double dmu[n];
RD_UNROLL
    for (int k = 0; k < n; k++) dmu[k] = st.mus[k] - mup[k];
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double s = muf[i];
RD_UNROLL
      for (int k = 0; k < n; k++) s += T[k * n + i] * dmu[k];
      st.mus[i] = s;
    }
  }
  if (DO_SIM) {                              // KalmanTV.h:511-527
    double dx[n], mt[n], V[NS], Sf[NS], zt[n];
RD_UNROLL
    for (int k = 0; k < n; k++) dx[k] = st.x[k] - mup[k];
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double s = muf[i];
RD_UNROLL
      for (int k = 0; k < n; k++) s += T[k * n + i] * dx[k];
      mt[i] = s;
    }
RD_UNROLL
    for (int i = 0; i < NS; i++) {
      Sf[i] = rec2[(n + i) * kTile];
      V[i] = Sf[i];
    }
    // V~ = Sigma_f - T~^T (Q Sigma_f): row k of Q Sigma_f is recomputed (T was overwritten)
RD_UNROLL
    for (int k = 0; k < n; k++) {
      double Tk[n];
RD_UNROLL
      for (int j = 0; j < n; j++) {
        double s = 0.0;
RD_UNROLL
        for (int l = 0; l < n; l++) s += P.Q[k * n + l] * Sf[sidx<n>(l, j)];
        Tk[j] = s;
      }
RD_UNROLL
      for (int i = 0; i < n; i++)
RD_UNROLL
        for (int j = i; j < n; j++) V[sidx<n>(i, j)] -= T[k * n + i] * Tk[j];
    }
RD_UNROLL
    for (int i = 0; i < n; i++) zt[i] = zt_ptr[i];
    int f2 = state_sim<n>(st.x, mt, V, zt);
    if (f2 && !st.fail) st.fail = f2;
  }
  if (DO_VAR) {                              // Sigma_s = Sigma_f + T~^T D T~  (:461-465)
    double Sn[NS];
RD_UNROLL
    for (int i = 0; i < n; i++) {
      double G[n];
RD_UNROLL
      for (int l = 0; l < n; l++) {
        double s = 0.0;
RD_UNROLL
        for (int k = 0; k < n; k++) s += T[k * n + i] * st.Ss[sidx<n>(k, l)];
        G[l] = s;
      }
RD_UNROLL
      for (int j = i; j < n; j++) {
        double s = rec2[(n + sidx<n>(i, j)) * kTile];
RD_UNROLL
        for (int l = 0; l < n; l++) s += G[l] * T[l * n + j];
        Sn[sidx<n>(i, j)] = s;
      }
    }
RD_UNROLL
    for (int i = 0; i < NS; i++) st.Ss[i] = Sn[i];
  }
}

// Generic backward pass: plain global loads and direct (per-thread) global stores.  Used by
// the rolled family, by every log-likelihood kernel and by tests/hostsim; the register family's
// solve_mv / solve_sim / solve kernels use the shared-memory staged version in
// kalman_staged.cuh.  DO_MEAN / DO_VAR: smooth_mv; DO_SIM: smooth_sim; LOGLIK: no per-step
// output, accumulate the thinned log-likelihood of the mean (DO_MEAN) or draw (DO_SIM).
template <class Ode, int n, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
RD_HD void backward_body(const Prior<n, Ode::n_meas> &P, const SolveArgs &a, long b) {
  constexpr int m = Ode::n_meas, NS = Sym<n>::size, NE = n + NS;
  const int N = P.n_eval;
  const long T1 = (long)N + 1;
  SmoothState<n> st;
  st.fail = 0;
  double lp = 0.0;
  int d = LOGLIK ? a.n_obs_times - 1 : -1;
  const double *zb = DO_SIM ? a.z + b * a.z_stride : nullptr;
  double *mu_o = (DO_MEAN && !LOGLIK) ? a.mu_out + b * T1 * n : nullptr;
  double *var_o = (DO_VAR && !LOGLIK) ? a.var_out + b * T1 * n * n : nullptr;
  double *x_o = (DO_SIM && !LOGLIK) ? a.x_out + b * T1 * n : nullptr;
  for (int t = N; t >= 1; t--) {
    const double *rec = a.hist + hist_offset(b, t - 1, N, NE);
    const double *zt = DO_SIM ? zb + (long)(t - 1) * n : nullptr;
    if (t == N)
      backward_terminal<n, DO_MEAN, DO_VAR, DO_SIM>(rec, zt, st);
    else
      backward_step<n, m, DO_MEAN, DO_VAR, DO_SIM>(P, rec, zt, st);
    if (!LOGLIK) {
      if (DO_MEAN) {
RD_UNROLL
        for (int i = 0; i < n; i++) mu_o[(long)t * n + i] = st.mus[i];
      }
      if (DO_SIM) {
RD_UNROLL
        for (int i = 0; i < n; i++) x_o[(long)t * n + i] = st.x[i];
      }
      if (DO_VAR) {
RD_UNROLL
        for (int j = 0; j < n; j++)
RD_UNROLL
          for (int i = 0; i < n; i++) var_o[((long)t * n + j) * n + i] = st.Ss[sidx<n>(i, j)];
      }
    } else {
      while (d >= 0 && a.thin_index[d] == t) {
        lp += loglik_terms<n>(a, d, DO_SIM ? st.x : st.mus);
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
  if (a.status && st.fail && a.status[b] == 0) a.status[b] = st.fail;
}
