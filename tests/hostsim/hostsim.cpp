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
static Prior<n, m> make_prior(const HostPrior &h) {
  Prior<n, m> P;
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
  return P;
}

template <class Ode, int n, bool UNROLL>
static int run(const HostPrior &h, int ig, int kind, SolveArgs a) {
  auto P = make_prior<n, Ode::n_meas>(h);
  std::vector<double> hist((size_t)h.n_eval * (n + Sym<n>::size) * ((a.B + 31) / 32 * 32));
  a.hist = hist.data();
  for (long b = 0; b < a.B; b++) {
    if (ig == 0) Impl<UNROLL>::template forward<Ode, n, IG_PROBDE>(P, a, b);
    else if (ig == 1) Impl<UNROLL>::template forward<Ode, n, IG_SCHOBERT>(P, a, b);
    else Impl<UNROLL>::template forward<Ode, n, IG_CHKREBTII>(P, a, b);
  }
  for (long b = 0; b < a.B; b++) {
    switch (kind) {
      case 0: Impl<UNROLL>::template backward<Ode, n, true, true, false, false>(P, a, b); break;
      case 1: Impl<UNROLL>::template backward<Ode, n, false, false, true, false>(P, a, b); break;
      case 2: Impl<UNROLL>::template backward<Ode, n, true, true, true, false>(P, a, b); break;
      case 3: Impl<UNROLL>::template backward<Ode, n, true, false, false, true>(P, a, b); break;
      case 4: Impl<UNROLL>::template backward<Ode, n, false, false, true, true>(P, a, b); break;
      default: return -1;
    }
  }
  return 0;
}

extern "C" int hostsim_solve(int ode, int n, int m, int n_eval, double tmin, double tmax, const double *W,
                             const double *Q, const double *c, const double *R, int ig, int kind,
                             int unroll, long B, const double *x0, const double *theta, const double *z,
                             long z_stride, double *x_out, double *mu_out, double *var_out, int *status,
                             const int *thin_index, int n_obs_times, const int *state_index,
                             int n_obs_comp, const double *Y, double gamma, double *loglik) {
  HostPrior h{n, m, n_eval, tmin, tmax, W, Q, c, R};
  SolveArgs a;
  memset(&a, 0, sizeof a);
  a.B = B;
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
#define CASE(ODE, NN, ID)                                                     \
  if (ode == ID && n == NN)                                                   \
    return unroll ? run<ODE, NN, true>(h, ig, kind, a) : run<ODE, NN, false>(h, ig, kind, a);
  CASE(OdeChkrebtii, 4, 0)
  CASE(OdeFitz, 6, 1)
  CASE(OdeLorenz63, 9, 2)
  CASE(OdeMseir, 15, 3)
  return -2;
}
