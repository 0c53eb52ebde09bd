// kalman_tps.cuh -- thread-per-system forward filter and backward smoother bodies.
//
// One thread runs one whole KalmanODE solve.  The bodies are __host__ __device__ so
// that tests can execute exactly this code on the CPU (tests/hostsim) before any GPU
// time is spent; the product only ever launches them as CUDA kernels (kernels.cu).
//
// Data layout in HBM
//   history  hist[tile][s][e][lane]  tile = b/32, lane = b%32, s = 0..N-1 (time index s+1),
//            e = 0..NE-1:  e < n: mu_filt element e;  e >= n: Sigma_filt packed upper e-n.
//            The 32 threads of a warp store 32 consecutive doubles per element (coalesced
//            256 B) and a warp's whole record for one time index is contiguous (one bulk
//            copy into shared memory in the backward pass).  Only the FILTERED moments are kept:
//            the predicted moments the smoother needs (mu_{t+1|t}, Sigma_{t+1|t}) are
//            recomputed from them in the backward pass with the same code the forward
//            pass ran, which gives bit-identical values and saves
//            (n + n^2) + (n^2 - n(n+1)/2) doubles per step of traffic versus the
//            reference's four dense history arrays (rodeo/cython/KalmanODE.pyx:196-203).
//   outputs  the reference's return layout per system (KalmanODE.pyx:394, 427, 463):
//            x_out, mu_out [b][t][i], var_out [b][t][j][i].
//
// Call order mirrors rodeo/cython/KalmanODE.pyx:293-332 (forward), :445-461 (solve_mv),
// :409-425 (solve_sim), :381-392 (solve).
#pragma once
#include "kalman_common.cuh"
#include "ode_functors.cuh"

namespace rodeo {

struct SolveArgs {
  long B;              // systems in this launch
  const double *x0;    // [B][n]
  const double *theta; // [B][n_theta] or nullptr
  const double *z;     // column-major (n x N) per system, or nullptr
  long z_stride;       // 0: shared, n*N: per system
  double *hist;        // [ceil(B/32)][R][NE][32], R = n_records(N, ck)
  int ck;              // checkpoint interval: records exist for time indices N, N-ck, N-2ck, ... >= 1
  int perm16;          // history position of system b is perm16(b) (kalman_common.cuh), else b
  int n_steps;         // N + 1 (time points per system in the outputs)
  double *x_out, *mu_out, *var_out;
  int *status;
  // thinned log-likelihood epilogue (examples/inference.py:54-56, 66-94)
  const int *thin_index;  // [n_obs_times] ascending grid indices
  const int *state_index; // [n_obs_comp]
  const double *Y;        // [n_obs_times][n_obs_comp]
  double *loglik;         // [B]
  const double *tab;      // shared-covariance tables (kalman_sc.inl), or nullptr
  int tab_fail;           // Cholesky failure code of the table pass (reported as every system's status)
  double gamma;
  int n_obs_times, n_obs_comp;
};

enum { IG_PROBDE = 0, IG_SCHOBERT = 1, IG_CHKREBTII = 2 };

}  // namespace rodeo

// RD_UNROLL: every loop; RD_UNROLL_IN: innermost accumulation loops of the math routines, which
// are unrolled in BOTH policies (in the rolled family that gives the local-memory loads of one
// dot product instruction-level parallelism instead of one dependent load per iteration).
#define RD_UNROLL _Pragma("unroll")
#define RD_UNROLL_IN _Pragma("unroll")
namespace rodeo {
namespace reg {
#include "kalman_math.inl"
#include "kalman_tps.inl"
#include "kalman_sc.inl"
}  // namespace reg
}  // namespace rodeo
#undef RD_UNROLL
#define RD_UNROLL _Pragma("unroll 1")
namespace rodeo {
namespace loc {
#include "kalman_math.inl"
#include "kalman_tps.inl"
}  // namespace loc
}  // namespace rodeo
#undef RD_UNROLL
#undef RD_UNROLL_IN

namespace rodeo {

// Family selector: Fam<UNROLL, BD, Ode, K> is the dense family with n_state = K (BD = false)
// or the block-diagonal family with block size K (BD = true), in the register (UNROLL) or
// rolled loop policy.
template <bool UNROLL, bool BD, class Ode, int K, bool ST = false>
struct FamSel;
template <class Ode, int K>
struct FamSel<true, false, Ode, K, false> { using type = reg::DenseFam<Ode, K>; };
template <class Ode, int K, bool ST>
struct FamSel<true, true, Ode, K, ST> { using type = reg::BdFam<Ode, K, ST>; };
template <class Ode, int K>
struct FamSel<false, false, Ode, K, false> { using type = loc::DenseFam<Ode, K>; };
template <class Ode, int K>
struct FamSel<false, true, Ode, K, false> { using type = loc::BdFam<Ode, K>; };

template <bool UNROLL>
struct Bodies;
template <>
struct Bodies<true> {
  template <class Fam, int IG>
  RD_HD static void forward(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
    reg::forward_body<Fam, IG>(P, a, b);
  }
  template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
  RD_HD static void backward(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
    reg::backward_body<Fam, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(P, a, b);
  }
};
template <>
struct Bodies<false> {
  template <class Fam, int IG>
  RD_HD static void forward(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
    loc::forward_body<Fam, IG>(P, a, b);
  }
  template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
  RD_HD static void backward(const typename Fam::PriorT &P, const SolveArgs &a, long b) {
    loc::backward_body<Fam, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(P, a, b);
  }
};

}  // namespace rodeo
