// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the very same thread-per-system kernel bodies the CUDA kernels launch
// (rodeo_legacy_b200/csrc/kalman_tps.cuh, __host__ __device__) on the CPU, one "thread"
// after another, so indexing / layout / algebra bugs are caught without a GPU.  It is not
// part of the product library and the product never loads it.
#include <cstring>
#include <vector>

#include "../../rodeo_legacy_b200/csrc/kalman_tps.cuh"

using namespace rodeo;

struct HostPrior {
  int n, m, n_eval;
  double tmin, tmax;
  const double *W, *Q, *c, *R;
};

template <int n, int m>
static void fill_prior(Prior<n, m> &P, const HostPrior &h) {
  for (int i = 0; i < n; i++) {
    P.c[i] = h.c[i];
    for (int j = 0; j < n; j++) {
      P.Q[i * n + j] = h.Q[i + (size_t)j * n];
      if (i <= j) P.R[sidx<n>(i, j)] = h.R[i + (size_t)j * n];
    }
  }
  for (int a = 0; a < m; a++)
    for (int j = 0; j < n; j++) P.W[a * n + j] = h.W[a + (size_t)j * m];
  P.tmin = h.tmin;
  P.tmax = h.tmax;
  P.n_eval = h.n_eval;
}

template <int p, int nb>
static void fill_prior(PriorBD<p, nb> &P, const HostPrior &h) {
  const int n = p * nb;
  for (int b = 0; b < nb; b++)
    for (int i = 0; i < p; i++) {
      P.w[b][i] = h.W[b + (size_t)(b * p + i) * nb];
      for (int j = 0; j < p; j++) {
        P.Q[b][i * p + j] = h.Q[(b * p + i) + (size_t)(b * p + j) * n];
        if (i <= j) P.R[b][sidx<p>(i, j)] = h.R[(b * p + i) + (size_t)(b * p + j) * n];
      }
    }
  for (int i = 0; i < n; i++) P.c[i] = h.c[i];
  P.tmin = h.tmin;
  P.tmax = h.tmax;
  P.n_eval = h.n_eval;
}

template <class Fam, bool UNROLL, int CK>
static int run_ck(const typename Fam::PriorT &P, int kind, const SolveArgs &a) {
  for (long b = 0; b < a.B; b++) {
    switch (kind) {
      case 0: Bodies<UNROLL>::template backward<Fam, CK, true, true, false, false>(P, a, b); break;
      case 1: Bodies<UNROLL>::template backward<Fam, CK, false, false, true, false>(P, a, b); break;
      case 2: Bodies<UNROLL>::template backward<Fam, CK, true, true, true, false>(P, a, b); break;
      case 3: Bodies<UNROLL>::template backward<Fam, CK, true, false, false, true>(P, a, b); break;
      case 4: Bodies<UNROLL>::template backward<Fam, CK, false, false, true, true>(P, a, b); break;
      default: return -1;
    }
  }
  return 0;
}

// shared-covariance mode: tables once, then the mean-only bodies (register namespace only)
template <class Fam>
static int run_sc(const HostPrior &h, int ig, int kind, SolveArgs a, double *var_single) {
  typename Fam::PriorT P;
  fill_prior(P, h);
  using T = reg::ScTab<Fam>;
  const int N = h.n_eval, n = Fam::n;
  std::vector<double> tab((size_t)T::total(N), 0.0);
  int fail = 0;
  for (int u = 0; u < Fam::kBlocks; u++) {
    int f = reg::sc_tables_unit<Fam>(P, ig == 1, tab.data(), u);
    if (f && !fail) fail = f;
  }
  if (var_single) memcpy(var_single, tab.data() + T::offVar(N), sizeof(double) * (N + 1) * n * n);
  std::vector<double> hist((size_t)N * n * ((a.B + 31) / 32 * 32));
  a.hist = hist.data();
  a.tab = tab.data();
  a.tab_fail = fail;
  for (long b = 0; b < a.B; b++) reg::sc_forward_body<Fam>(P, a, b);
  for (long b = 0; b < a.B; b++) {
    switch (kind) {
      case 0: reg::sc_backward_body<Fam, true, false, false>(P, a, b); break;
      case 1: reg::sc_backward_body<Fam, false, true, false>(P, a, b); break;
      case 2: reg::sc_backward_body<Fam, true, true, false>(P, a, b); break;
      case 3: reg::sc_backward_body<Fam, true, false, true>(P, a, b); break;
      case 4: reg::sc_backward_body<Fam, false, true, true>(P, a, b); break;
      default: return -1;
    }
  }
  return 0;
}

template <class Fam, bool UNROLL>
static int run(const HostPrior &h, int ig, int kind, SolveArgs a) {
  typename Fam::PriorT P;
  fill_prior(P, h);
  std::vector<double> hist((size_t)n_records(h.n_eval, a.ck) * Fam::NE * ((a.B + 31) / 32 * 32));
  a.hist = hist.data();
  for (long b = 0; b < a.B; b++) {
    if (ig == 0) Bodies<UNROLL>::template forward<Fam, IG_PROBDE>(P, a, b);
    else if (ig == 1) Bodies<UNROLL>::template forward<Fam, IG_SCHOBERT>(P, a, b);
    else Bodies<UNROLL>::template forward<Fam, IG_CHKREBTII>(P, a, b);
  }
  if (a.ck == 1) return run_ck<Fam, UNROLL, 1>(P, kind, a);
  if (ig != 0) return -3;
  if (a.ck == 2) return run_ck<Fam, UNROLL, 2>(P, kind, a);
  if (a.ck == 3) return run_ck<Fam, UNROLL, 3>(P, kind, a);
  if (a.ck == 4) return run_ck<Fam, UNROLL, 4>(P, kind, a);
  return -4;
}

extern "C" int hostsim_solve(int ode, int n, int m, int n_eval, double tmin, double tmax, const double *W,
                             const double *Q, const double *c, const double *R, int ig, int kind,
                             int unroll, int blockdiag, int ck, int shared_cov, long B, const double *x0, const double *theta, const double *z,
                             long z_stride, double *x_out, double *mu_out, double *var_out, int *status,
                             const int *thin_index, int n_obs_times, const int *state_index,
                             int n_obs_comp, const double *Y, double gamma, double *loglik) {
  HostPrior h{n, m, n_eval, tmin, tmax, W, Q, c, R};
  SolveArgs a;
  memset(&a, 0, sizeof a);
  a.B = B;
  a.ck = ck;
  a.n_steps = n_eval + 1;
  a.x0 = x0;
  a.theta = theta;
  a.z = z;
  a.z_stride = z_stride;
  a.x_out = x_out;
  a.mu_out = mu_out;
  a.var_out = var_out;
  a.status = status;
  a.thin_index = thin_index;
  a.n_obs_times = n_obs_times;
  a.state_index = state_index;
  a.n_obs_comp = n_obs_comp;
  a.Y = Y;
  a.gamma = gamma;
  a.loglik = loglik;
#define CASE(ODE, NN, PP, ID)                                                             \
  if (ode == ID && n == NN && shared_cov) {                                               \
    if (ig == 2) return -5;                                                               \
    return blockdiag ? run_sc<reg::BdFam<ODE, PP>>(h, ig, kind, a, var_out)               \
                     : run_sc<reg::DenseFam<ODE, NN>>(h, ig, kind, a, var_out);           \
  }                                                                                       \
  if (ode == ID && n == NN) {                                                             \
    if (blockdiag == 2) /* structured: upper-triangular Q blocks, w = e_order */          \
      return run<reg::BdFam<ODE, PP, true>, true>(h, ig, kind, a);                        \
    if (blockdiag)                                                                        \
      return unroll ? run<reg::BdFam<ODE, PP>, true>(h, ig, kind, a)                      \
                    : run<loc::BdFam<ODE, PP>, false>(h, ig, kind, a);                    \
    return unroll ? run<reg::DenseFam<ODE, NN>, true>(h, ig, kind, a)                     \
                  : run<loc::DenseFam<ODE, NN>, false>(h, ig, kind, a);                   \
  }
  CASE(OdeChkrebtii, 4, 4, 0)
  CASE(OdeFitz, 6, 3, 1)
  CASE(OdeLorenz63, 9, 3, 2)
  CASE(OdeMseir, 15, 3, 3)
  return -2;
}
