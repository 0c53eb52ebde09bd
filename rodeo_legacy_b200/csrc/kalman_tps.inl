// kalman_tps.inl -- thread-per-system forward / backward bodies (see kalman_tps.cuh), for the
// two problem families:
//   DenseFam<Ode, n>   general dense Q, R, W                     (record NE = n + n(n+1)/2)
//   BdFam<Ode, p>      block-diagonal Q, R with nb = n_meas blocks of size p, W row a inside
//                      block a (kalman_common.cuh)               (record NE = n + nb p(p+1)/2)
// A family provides: step<IG>() -- one filter step t -> t+1 on (mu, Sigma) held by the thread;
// forward<IG>() -- the whole forward filter of one system, storing every CK-th filtered record;
// bw_terminal / bw_step -- the smoother recursion on a record read with a given stride;
// var_at -- element (i, j) of the carried smoothed covariance.
//
// Checkpointing (SolveArgs::ck = CK): only the filtered moments at time indices t with
// (N - t) % CK == 0 are written to HBM (record r = (N - t) / CK).  The backward pass re-runs the
// filter from each checkpoint for the CK-1 time indices above it -- the same code on the same
// inputs, hence bit-identical moments -- trading cheap FP64 work for CK-fold less history
// traffic.  CK = 1 stores everything.  CK > 1 is offered with the probde interrogation only.

template <class Ode_, int n_>
struct DenseFam {
  using Ode = Ode_;
  static constexpr int n = n_, m = Ode::n_meas, NS = Sym<n_>::size, NE = n_ + Sym<n_>::size;
  static constexpr bool kBlockDiag = false;
  static constexpr int kBlocks = 1, kBlockSize = n_;
  using PriorT = Prior<n_, Ode_::n_meas>;

  // One filter step, time index t -> t+1 (KalmanODE.pyx:299-332): predict, interrogate, update.
  template <int IG>
  RD_HD static int step(const PriorT &P, int t, const double *th, const double *zb, double *mu, double *S) {
    double mup[n], Sp[NS], y[m];
    int fail = 0;
    predict_mean<n>(P.Q, P.c, mu, mup);
    predict_var<n, false>(P.Q, P.R, S, Sp, nullptr);
    const double tt = P.tmin + (P.tmax - P.tmin) * (t + 1) / P.n_eval;
    if (IG == IG_CHKREBTII) {  // x_state ~ N(mu_pred, Sigma_pred) with z[:, t]  (KalmanODE.pyx:13-49)
      double xs[n], V[NS], zt[n];
RD_UNROLL
      for (int i = 0; i < NS; i++) V[i] = Sp[i];
RD_UNROLL
      for (int i = 0; i < n; i++) zt[i] = zb[(long)t * n + i];
      fail = state_sim<n>(xs, mup, V, zt);
      Ode::template eval<n>(xs, tt, th, y);
    } else {
      Ode::template eval<n>(mup, tt, th, y);
    }
    int f = update_dense<n, m>(P.W, mup, Sp, y, IG == IG_SCHOBERT);
    if (f && !fail) fail = f;
RD_UNROLL
    for (int i = 0; i < n; i++) mu[i] = mup[i];
RD_UNROLL
    for (int i = 0; i < NS; i++) S[i] = Sp[i];
    return fail;
  }

  template <int STRIDE, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
  RD_HD static int bw_terminal(const PriorT &, const double *rec, const double *zt, double *mus, double *Ss,
                               double *x) {
    return smooth_terminal<n, STRIDE, DO_MEAN, DO_VAR, DO_SIM>(rec, rec + n * STRIDE, zt, mus, Ss, x);
  }
  template <int STRIDE, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
  RD_HD static int bw_step(const PriorT &P, const double *rec, const double *zt, double *mus, double *Ss,
                           double *x) {
    return smooth_step<n, STRIDE, DO_MEAN, DO_VAR, DO_SIM>(P.Q, P.c, P.R, rec, rec + n * STRIDE, zt, mus, Ss, x);
  }
  RD_HD static double var_at(const double *Ss, int i, int j) { return Ss[sidx<n>(i, j)]; }
};

// ST ("structured"): every block's Q is upper triangular and w = e_{Ode::order} (ibm_init + zero_pad priors of
// a first/second-order ODE: every BASELINE config); see kalman_math.inl.  Selected on the host; same results.
template <class Ode_, int p_, bool ST_ = false>
struct BdFam {
  using Ode = Ode_;
  static constexpr bool kUT = ST_;
  static constexpr int kWK = ST_ ? Ode_::order : -1;
  using General = BdFam<Ode_, p_, false>;        // the same family without the specialisation (same results)
  static_assert(!ST_ || Ode_::order < p_, "the unit entry of w must lie inside the block");
  static constexpr int p = p_, nb = Ode::n_meas, n = p_ * Ode_::n_meas, m = Ode::n_meas;
  static constexpr int PS = Sym<p_>::size, NS = Ode_::n_meas * Sym<p_>::size, NE = n + NS;
  static constexpr bool kBlockDiag = true;
  static constexpr int kBlocks = Ode_::n_meas, kBlockSize = p_;
  using PriorT = PriorBD<p_, Ode_::n_meas>;

  template <int IG>
  RD_HD static int step(const PriorT &P, int t, const double *th, const double *zb, double *mu, double *S) {
    double mup[n], Sp[NS], y[m];
    int fail = 0;
RD_UNROLL
    for (int blk = 0; blk < nb; blk++) {
      predict_mean<p, kUT>(P.Q[blk], P.c + blk * p, mu + blk * p, mup + blk * p);
      predict_var<p, false, kUT>(P.Q[blk], P.R[blk], S + blk * PS, Sp + blk * PS, nullptr);
    }
    const double tt = P.tmin + (P.tmax - P.tmin) * (t + 1) / P.n_eval;
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
RD_UNROLL
    for (int blk = 0; blk < nb; blk++) {
      int f = update_block<p, kWK>(P.w[blk], mup + blk * p, Sp + blk * PS, y[blk], IG == IG_SCHOBERT);
      if (f && !fail) fail = blk + 1;
    }
RD_UNROLL
    for (int i = 0; i < n; i++) mu[i] = mup[i];
RD_UNROLL
    for (int i = 0; i < NS; i++) S[i] = Sp[i];
    return fail;
  }

  template <int STRIDE, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
  RD_HD static int bw_terminal(const PriorT &, const double *rec, const double *zt, double *mus, double *Ss,
                               double *x) {
    int fail = 0;
RD_UNROLL
    for (int blk = 0; blk < nb; blk++) {
      int f = smooth_terminal<p, STRIDE, DO_MEAN, DO_VAR, DO_SIM>(
          rec + blk * p * STRIDE, rec + (n + blk * PS) * STRIDE, DO_SIM ? zt + blk * p : nullptr, mus + blk * p,
          Ss + blk * PS, x + blk * p);
      if (f && !fail) fail = blk * p + f;
    }
    return fail;
  }
  template <int STRIDE, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
  RD_HD static int bw_step(const PriorT &P, const double *rec, const double *zt, double *mus, double *Ss,
                           double *x) {
    int fail = 0;
RD_UNROLL
    for (int blk = 0; blk < nb; blk++) {
      int f = smooth_step<p, STRIDE, DO_MEAN, DO_VAR, DO_SIM, true, kUT>(
          P.Q[blk], P.c + blk * p, P.R[blk], rec + blk * p * STRIDE, rec + (n + blk * PS) * STRIDE,
          DO_SIM ? zt + blk * p : nullptr, mus + blk * p, Ss + blk * PS, x + blk * p);
      if (f && !fail) fail = blk * p + f;
    }
    return fail;
  }
  RD_HD static double var_at(const double *Ss, int i, int j) {
    return (i / p == j / p) ? Ss[(i / p) * PS + sidx<p>(i % p, j % p)] : 0.0;
  }
};

// Whole forward filter of one system; stores the filtered record of every time index t with
// (N - t) % ck == 0 as record r = (N - t) / ck.
template <class Fam, int IG>
RD_HD void forward_body(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
  using Ode = typename Fam::Ode;
  constexpr int n = Fam::n, NS = Fam::NS, NE = Fam::NE;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  double mu[n], S[NS], th[NT];
RD_UNROLL
  for (int i = 0; i < n; i++) mu[i] = a.x0[b * n + i];
RD_UNROLL
  for (int i = 0; i < NS; i++) S[i] = 0.0;
RD_UNROLL
  for (int i = 0; i < NT; i++) th[i] = (Ode::n_theta > 0) ? a.theta[b * Ode::n_theta + i] : 0.0;
  const int N = P.n_eval, ck = a.ck, R = n_records(N, ck);
  const long hpos = a.perm16 ? perm16(b) : b;     // position of this system in the history
  const double *zb = (IG == IG_CHKREBTII) ? a.z + b * a.z_stride : nullptr;
  int fail = 0;
  // time index t+1 is kept when its distance N-(t+1) from the end is a multiple of ck: records r = R-1 down to 0.
  // A countdown and a running pointer: no division by the run-time interval inside the loop.
  int until = (N - 1) % ck;
  double *h = a.hist + hist_offset(hpos, R - 1, R, NE);
  for (int t = 0; t < N; t++) {
    int f = Fam::template step<IG>(P, t, th, zb, mu, S);
    if (f && !fail) fail = f;
    if (until == 0) {
RD_UNROLL
      for (int i = 0; i < n; i++) h[i * kTile] = mu[i];
RD_UNROLL
      for (int i = 0; i < NS; i++) h[(n + i) * kTile] = S[i];
      h -= NE * kTile;
      until = ck;
    }
    until--;
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

// Deliver the smoothed quantities of time index t: per-step outputs in the reference's return
// layout, or the thinned log-likelihood accumulation.
template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
RD_HD void emit_time(const SolveArgs &a, long b, int t, const double *mus, const double *Ss, const double *x,
                     double &lp, int &d) {
  constexpr int n = Fam::n;
  if (!LOGLIK) {
    const long T1 = (long)a.n_steps;
    if (DO_MEAN) {
      double *o = a.mu_out + (b * T1 + t) * n;
RD_UNROLL
      for (int i = 0; i < n; i++) o[i] = mus[i];
    }
    if (DO_SIM) {
      double *o = a.x_out + (b * T1 + t) * n;
RD_UNROLL
      for (int i = 0; i < n; i++) o[i] = x[i];
    }
    if (DO_VAR) {
      double *o = a.var_out + (b * T1 + t) * n * n;
RD_UNROLL
      for (int j = 0; j < n; j++)
RD_UNROLL
        for (int i = 0; i < n; i++) o[j * n + i] = Fam::var_at(Ss, i, j);
    }
  } else {
    while (d >= 0 && a.thin_index[d] == t) {
      lp += loglik_terms<n>(a, d, DO_SIM ? x : mus);
      d--;
    }
  }
}

// Generic backward pass: plain global loads and direct (per-thread) global stores.  Used by
// the rolled family, by every log-likelihood kernel and by tests/hostsim; the register
// families' solve_mv / solve_sim / solve kernels use the shared-memory staged versions
// (kalman_staged.cuh, kalman_bdw.cuh).  DO_MEAN / DO_VAR: smooth_mv; DO_SIM: smooth_sim;
// LOGLIK: no per-step output, accumulate the thinned log-likelihood of the mean (DO_MEAN) or
// draw (DO_SIM).  CK: checkpoint interval the forward pass used (a.ck == CK).
template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
RD_HD void backward_body(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
  using Ode = typename Fam::Ode;
  constexpr int n = Fam::n, NS = Fam::NS, NE = Fam::NE;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  const int N = P.n_eval, R = n_records(N, CK);
  const long hpos = a.perm16 ? perm16(b) : b;     // position of this system in the history
  double mus[n], Ss[NS], x[n], th[NT];
  int fail = 0;
  double lp = 0.0;
  int d = LOGLIK ? a.n_obs_times - 1 : -1;
  const double *zb = DO_SIM ? a.z + b * a.z_stride : nullptr;
  if (CK > 1) {
RD_UNROLL
    for (int i = 0; i < NT; i++) th[i] = (Ode::n_theta > 0) ? a.theta[b * Ode::n_theta + i] : 0.0;
  }
  {  // terminal time index N = record 0
    const double *rec = a.hist + hist_offset(hpos, 0, R, NE);
    fail = Fam::template bw_terminal<kTile, DO_MEAN, DO_VAR, DO_SIM>(P, rec, DO_SIM ? zb + (long)(N - 1) * n : nullptr,
                                                                     mus, Ss, x);
    emit_time<Fam, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(a, b, N, mus, Ss, x, lp, d);
  }
  int r = 1;
  for (int t_hi = N; t_hi > 1;) {   // times t_hi-1 .. max(t_hi-CK, 1) are smoothed in this group
    const int base = t_hi - CK;      // checkpointed time index of the group, if >= 1
    if (CK == 1) {
      const double *rec = a.hist + hist_offset(hpos, r, R, NE);
      int f = Fam::template bw_step<kTile, DO_MEAN, DO_VAR, DO_SIM>(P, rec, DO_SIM ? zb + (long)(base - 1) * n : nullptr,
                                                                    mus, Ss, x);
      if (f && !fail) fail = f;
      emit_time<Fam, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(a, b, base, mus, Ss, x, lp, d);
    } else {
      // start of the re-run: the checkpoint (time base >= 1) or the known initial state (time 0)
      constexpr int NSCR = CK > 1 ? (CK - 1) * NE : 1;
      double cur[NE], scratch[NSCR];
      const int t0 = base >= 1 ? base : 0;
      const int nfwd = t_hi - 1 - t0;  // recomputed time indices t0+1 .. t_hi-1
      if (base >= 1) {
        const double *rec = a.hist + hist_offset(hpos, r, R, NE);
RD_UNROLL
        for (int i = 0; i < NE; i++) cur[i] = rec[i * kTile];
      } else {
RD_UNROLL
        for (int i = 0; i < n; i++) cur[i] = a.x0[b * n + i];
RD_UNROLL
        for (int i = 0; i < NS; i++) cur[n + i] = 0.0;
      }
      {
        double fm[n], fS[NS];
RD_UNROLL
        for (int i = 0; i < n; i++) fm[i] = cur[i];
RD_UNROLL
        for (int i = 0; i < NS; i++) fS[i] = cur[n + i];
RD_UNROLL
        for (int j = 0; j < CK - 1; j++) {
          if (j < nfwd) {
            Fam::template step<IG_PROBDE>(P, t0 + j, th, nullptr, fm, fS);
RD_UNROLL
            for (int i = 0; i < n; i++) scratch[j * NE + i] = fm[i];
RD_UNROLL
            for (int i = 0; i < NS; i++) scratch[j * NE + n + i] = fS[i];
          }
        }
      }
RD_UNROLL
      for (int j = CK - 2; j >= 0; j--) {
        if (j < nfwd) {
          const int t = t0 + j + 1;
          int f = Fam::template bw_step<1, DO_MEAN, DO_VAR, DO_SIM>(P, scratch + j * NE,
                                                                    DO_SIM ? zb + (long)(t - 1) * n : nullptr, mus, Ss, x);
          if (f && !fail) fail = f;
          emit_time<Fam, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(a, b, t, mus, Ss, x, lp, d);
        }
      }
      if (base >= 1) {
        int f = Fam::template bw_step<1, DO_MEAN, DO_VAR, DO_SIM>(P, cur, DO_SIM ? zb + (long)(base - 1) * n : nullptr,
                                                                  mus, Ss, x);
        if (f && !fail) fail = f;
        emit_time<Fam, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(a, b, base, mus, Ss, x, lp, d);
      }
    }
    t_hi = base >= 1 ? base : 1;
    r++;
  }
  // time index 0: the initial state is known (KalmanODE.pyx:445, 409); Sigma_0 = 0
  {
    double x0v[n], S0[NS];
RD_UNROLL
    for (int i = 0; i < n; i++) x0v[i] = a.x0[b * n + i];
RD_UNROLL
    for (int i = 0; i < NS; i++) S0[i] = 0.0;
    emit_time<Fam, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(a, b, 0, x0v, S0, x0v, lp, d);
  }
  if (LOGLIK) {
    // constant of the normal density: -(log sqrt(2 pi) + log gamma) per observation
    const double cst = -0.91893853320467274178 - log(a.gamma);
    a.loglik[b] = lp + cst * (double)(a.n_obs_times * a.n_obs_comp);
  }
  if (a.status && fail && a.status[b] == 0) a.status[b] = fail;
}
