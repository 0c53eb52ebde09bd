// kalman_sc.cuh -- CUDA kernels of the shared-covariance mode (bodies: kalman_sc.inl).
//   kalman_sc_tables     one thread per diagonal block (or one thread, dense): covariance
//                        recursion -> gain / Cholesky tables + the (N+1, n, n) variance output
//   kalman_sc_forward    thread per system, mean-only filter, streams mu_filt (n doubles / step)
//   kalman_sc_backward   warp per 32-system tile, mean-only smoother / sampler; four records in flight
//                        in a TMA ring, outputs stored as sector-aligned 4-step blocks (BlockedRuns)
//   kalman_sc_loglik     mean-only smoother accumulating the thinned log-likelihood
#pragma once
#include "kalman_staged.cuh"

namespace rodeo {

constexpr int kThreadsPerBlockSc = 128;

template <class Fam>
__global__ void kalman_sc_tables(const __grid_constant__ typename Fam::PriorT P, int zero_H, double *tab,
                                 int *fail_out) {
  const int u = threadIdx.x;
  if (u < Fam::kBlocks) {
    const int f = reg::sc_tables_unit<Fam>(P, zero_H != 0, tab, u);
    if (f) atomicCAS(fail_out, 0, u * Fam::kBlockSize + f);
  }
}

template <class Fam>
__global__ void __launch_bounds__(kThreadsPerBlockSc)
kalman_sc_forward(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) reg::sc_forward_body<Fam>(P, a, b);
}

template <class Fam, bool DO_MEAN, bool DO_SIM>
__global__ void __launch_bounds__(kThreadsPerBlockSc)
kalman_sc_loglik(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) reg::sc_backward_body<Fam, DO_MEAN, DO_SIM, true>(P, a, b);
}

constexpr int kScWarps = 4;
constexpr int kScBlock = 4;   // time indices per stored piece (BlockedRuns)
constexpr int kScRing = 4;    // filtered-mean records in flight per warp (TMA ring)

// Mean-only smoother / sampler.  The per-step arithmetic is tiny (two p x p mat-vecs per block), so the kernel
// lives on memory-level parallelism: each warp keeps kScRing records (n*32 doubles, one cp.async.bulk each) in
// flight in a shared-memory ring, and writes its outputs as 4-step blocks (BlockedRuns).
template <class Fam, bool DO_MEAN, bool DO_SIM>
__global__ void __launch_bounds__(kScWarps * 32)
kalman_sc_backward(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  constexpr int n = Fam::n, REC = n * kTile;
  using Runs = BlockedRuns<n, kScBlock, 32>;
  constexpr int NARR = (DO_MEAN ? 1 : 0) + (DO_SIM ? 1 : 0);
  constexpr int WARP_DOUBLES = (kScRing * REC + NARR * Runs::DOUBLES + kScRing + 15) / 16 * 16;
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long tile = (long)blockIdx.x * kScWarps + warp;
  if (tile * kTile >= a.B) return;
  double *ring = smem + (size_t)warp * WARP_DOUBLES;
  double *stage = ring + kScRing * REC;
  uint64_t *bar = reinterpret_cast<uint64_t *>(stage + NARR * Runs::DOUBLES);
  const int N = P.n_eval;
  const long T1 = (long)N + 1, b0 = tile * kTile, b = b0 + lane;
  const bool valid = b < a.B;
  const long bs = valid ? b : b0;
  const double *zb = DO_SIM ? a.z + bs * a.z_stride : nullptr;
  const double *hist_tile = a.hist + (size_t)tile * N * REC;     // record r = time index N - r
  constexpr uint32_t kRecBytes = REC * sizeof(double);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kScRing; k++) mbar_init(&bar[k], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kScRing; k++)
      if (k < N) {
        mbar_arrive_expect_tx(&bar[k], kRecBytes);
        bulk_g2s(ring + k * REC, hist_tile + (size_t)k * REC, kRecBytes, &bar[k]);
      }
  }
  Runs rm, rx;
  if (DO_MEAN) rm.init(stage, a.mu_out, b0, a.B, T1, lane);
  if (DO_SIM) rx.init(stage + (DO_MEAN ? Runs::DOUBLES : 0), a.x_out, b0, a.B, T1, lane);
  double mus[n], x[n];
  uint32_t phase = 0;
  for (int t = N; t >= 0; t--) {
    if (t >= 1) {
      const int r = N - t, k = r % kScRing;
      mbar_wait(&bar[k], (phase >> k) & 1u);
      phase ^= 1u << k;
      double cur[n];
#pragma unroll
      for (int i = 0; i < n; i++) cur[i] = ring[k * REC + i * kTile + lane];
      fence_proxy_async();   // this lane's reads of the slot are complete ...
      __syncwarp();          // ... for every lane, before lane 0 lets the TMA overwrite it
      if (lane == 0 && r + kScRing < N) {
        mbar_arrive_expect_tx(&bar[k], kRecBytes);
        bulk_g2s(ring + k * REC, hist_tile + (size_t)(r + kScRing) * REC, kRecBytes, &bar[k]);
      }
      const double *zt = DO_SIM ? zb + (long)(t - 1) * n : nullptr;
      if (t == N)
        reg::sc_backward_terminal<Fam, 1, DO_MEAN, DO_SIM>(P, a, cur, zt, mus, x);
      else
        reg::sc_backward_step<Fam, 1, DO_MEAN, DO_SIM>(P, a, t, cur, zt, mus, x);
    } else {
#pragma unroll
      for (int i = 0; i < n; i++) {
        const double v = valid ? a.x0[b * n + i] : 0.0;
        mus[i] = v;
        x[i] = v;
      }
    }
    if (DO_MEAN) rm.put(lane, t, mus);
    if (DO_SIM) rx.put(lane, t, x);
    __syncwarp();
    if (DO_MEAN) rm.store(t, lane);
    if (DO_SIM) rx.store(t, lane);
    __syncwarp();
  }
}

template <class Fam, bool DO_MEAN, bool DO_SIM>
inline size_t sc_backward_smem() {
  constexpr int NARR = (DO_MEAN ? 1 : 0) + (DO_SIM ? 1 : 0);
  constexpr int WARP_DOUBLES = (kScRing * Fam::n * kTile + NARR * BlockedRuns<Fam::n, kScBlock, 32>::DOUBLES + kScRing + 15) / 16 * 16;
  return (size_t)kScWarps * WARP_DOUBLES * sizeof(double);
}

// ---- launchers --------------------------------------------------------------------------------
template <class Fam>
cudaError_t launch_sc_tables(const HostPrior &h, int zero_H, double *tab, int *fail_out, cudaStream_t s) {
  kalman_sc_tables<Fam><<<1, 32, 0, s>>>(make_prior<Fam>(h), zero_H, tab, fail_out);
  return cudaGetLastError();
}
template <class Fam>
long sc_table_doubles(int N) {
  return reg::ScTab<Fam>::total(N);
}
template <class Fam>
long sc_table_var_offset(int N) {
  return reg::ScTab<Fam>::offVar(N);
}
template <class Fam>
cudaError_t launch_sc_forward(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  const unsigned grid = (unsigned)((a.B + kThreadsPerBlockSc - 1) / kThreadsPerBlockSc);
  kalman_sc_forward<Fam><<<grid, kThreadsPerBlockSc, 0, s>>>(make_prior<Fam>(h), a);
  return cudaGetLastError();
}
template <class Fam, bool DO_MEAN, bool DO_SIM, bool LOGLIK>
cudaError_t launch_sc_backward(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  if constexpr (LOGLIK) {
    const unsigned grid = (unsigned)((a.B + kThreadsPerBlockSc - 1) / kThreadsPerBlockSc);
    kalman_sc_loglik<Fam, DO_MEAN, DO_SIM><<<grid, kThreadsPerBlockSc, 0, s>>>(make_prior<Fam>(h), a);
  } else {
    const long n_tiles = (a.B + kTile - 1) / kTile;
    const unsigned grid = (unsigned)((n_tiles + kScWarps - 1) / kScWarps);
    auto kern = kalman_sc_backward<Fam, DO_MEAN, DO_SIM>;
    const size_t smem = sc_backward_smem<Fam, DO_MEAN, DO_SIM>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kScWarps * 32, smem, s>>>(make_prior<Fam>(h), a);
  }
  return cudaGetLastError();
}

}  // namespace rodeo
