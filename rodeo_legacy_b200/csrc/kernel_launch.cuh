// kernel_launch.cuh -- __global__ wrappers and launchers of the thread-per-system bodies.
#pragma once
#include "registry.cuh"

namespace rodeo {

constexpr int kThreadsPerBlock = 64;

template <int n, int m>
inline Prior<n, m> make_prior(const HostPrior &h) {
  Prior<n, m> P;
  for (int i = 0; i < n; i++) {
    P.c[i] = h.c[i];
    for (int j = 0; j < n; j++) {
      P.Q[i * n + j] = h.Q[i + (size_t)j * n];                 // column-major -> row-major
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

}  // namespace rodeo
#include "kalman_staged.cuh"
namespace rodeo {

template <class Ode, int n, int IG, bool UNROLL>
__global__ void __launch_bounds__(kThreadsPerBlock)
kalman_forward(const __grid_constant__ Prior<n, Ode::n_meas> P, const __grid_constant__ SolveArgs a) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) Impl<UNROLL>::template forward<Ode, n, IG>(P, a, b);
}

template <class Ode, int n, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK, bool UNROLL>
__global__ void __launch_bounds__(kThreadsPerBlock)
kalman_backward(const __grid_constant__ Prior<n, Ode::n_meas> P, const __grid_constant__ SolveArgs a) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) Impl<UNROLL>::template backward<Ode, n, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(P, a, b);
}

inline dim3 grid_for(long B) { return dim3((unsigned)((B + kThreadsPerBlock - 1) / kThreadsPerBlock)); }

template <class Ode, int n, int IG, bool UNROLL>
cudaError_t launch_forward(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  kalman_forward<Ode, n, IG, UNROLL><<<grid_for(a.B), kThreadsPerBlock, 0, s>>>(
      make_prior<n, Ode::n_meas>(h), a);
  return cudaGetLastError();
}

template <class Ode, int n, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK, bool UNROLL>
cudaError_t launch_backward(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  kalman_backward<Ode, n, DO_MEAN, DO_VAR, DO_SIM, LOGLIK, UNROLL>
      <<<grid_for(a.B), kThreadsPerBlock, 0, s>>>(make_prior<n, Ode::n_meas>(h), a);
  return cudaGetLastError();
}

template <class Ode, int n, bool UNROLL>
KernelSet make_kernel_set(const char *name) {
  KernelSet k;
  k.ode = Ode::id;
  k.n = n;
  k.m = Ode::n_meas;
  k.name = name;
  k.forward[IG_PROBDE] = launch_forward<Ode, n, IG_PROBDE, UNROLL>;
  k.forward[IG_SCHOBERT] = launch_forward<Ode, n, IG_SCHOBERT, UNROLL>;
  k.forward[IG_CHKREBTII] = launch_forward<Ode, n, IG_CHKREBTII, UNROLL>;
  if constexpr (UNROLL) {  // register family: history and outputs staged through shared memory
    k.backward[BK_MV] = launch_backward_staged<Ode, n, true, true, false>;
    k.backward[BK_SIM] = launch_backward_staged<Ode, n, false, false, true>;
    k.backward[BK_BOTH] = launch_backward_staged<Ode, n, true, true, true>;
  } else {
    k.backward[BK_MV] = launch_backward<Ode, n, true, true, false, false, UNROLL>;
    k.backward[BK_SIM] = launch_backward<Ode, n, false, false, true, false, UNROLL>;
    k.backward[BK_BOTH] = launch_backward<Ode, n, true, true, true, false, UNROLL>;
  }
  k.backward[BK_LL_MEAN] = launch_backward<Ode, n, true, false, false, true, UNROLL>;
  k.backward[BK_LL_DRAW] = launch_backward<Ode, n, false, false, true, true, UNROLL>;
  k.hist_elems = n + Sym<n>::size;
  return k;
}

}  // namespace rodeo
