// kalman_bdw.cuh -- backward (smoother) kernel of the block-diagonal family with one WARP PER
// DIAGONAL BLOCK: a CTA of nb = n_meas warps owns one tile of 32 systems; lane = system,
// warp = block.
//
// In the block-diagonal family the smoother decouples into nb independent p-dimensional
// smoothers (kalman_common.cuh), so the blocks of one system need not share a thread.  Spreading
// them over warps multiplies the resident parallelism by nb (MSEIR: 5x, Lorenz63: 3x) and cuts
// the registers per thread to one p x p block's worth, which is what the wide-state configs
// need: their batches (16,384 / 32,768 systems) are too small to fill 148 SMs with one thread
// per system.
//
// Data movement is the staged scheme of kalman_staged.cuh at CTA scope:
//   * history ring: ONE cp.async.bulk (TMA) per time index brings the tile's whole record
//     (NE*32 doubles, contiguous in HBM) into a two-slot shared-memory ring two time indices
//     ahead; completion on an mbarrier all warps wait on; each warp reads only its block's rows.
//   * checkpointed history (SolveArgs::ck > 1) is NOT offered by this kernel: re-running the
//     filter needs the blocks to exchange their predicted means through shared memory every
//     step, and on B200 the extra CTA barriers cost more than the saved HBM traffic
//     (measured: Lorenz63 5.7 -> 8.2 ms, MSEIR 17.9 -> 27.7 ms at interval 2).
//   * outputs: warps write their block's part of each system's output row into a staging tile
//     (two tiles, alternating per pass), then ALL threads store the tile so that every system's
//     contiguous run in out[b][t][...] is written by neighbouring threads (coalesced).  The
//     variance goes out in passes of PB blocks' columns so the tile stays ~10 KB for any n.
#pragma once
#include <cstdint>

#include "kalman_staged.cuh"

namespace rodeo {

template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
struct BdwLayout {
  static constexpr int p = Fam::p, nb = Fam::nb, n = Fam::n, NE = Fam::NE, PS = Fam::PS;
  static constexpr int REC = NE * kTile;
  static constexpr int VW = n * ((DO_MEAN ? 1 : 0) + (DO_SIM ? 1 : 0));   // "vector" pass width
  static constexpr int COLW = p * n;                                       // one block's variance columns
  static constexpr int KB = 4;                                             // time indices per stored mean / draw piece
  using Runs = BlockedRuns<n, KB, nb * 32>;
  static constexpr int NVEC = (DO_MEAN ? 1 : 0) + (DO_SIM ? 1 : 0);
  // mean / draw outputs as sector-aligned 4-step blocks (BlockedRuns) when the kernel also writes variances
  // (MSEIR solve_mv: 15.2 -> 14.2 ms); the draw-only kernels are register bound (two Cholesky factorisations
  // per block and step) and measured slower with it (Lorenz63 solve_sim 5.74 -> 6.01 ms): they keep the
  // one-step tile + 8-byte stores
  static constexpr bool kBlocked = DO_VAR;
  static constexpr int TW = (VW + 1) / 2 * 2;                              // one-step tile row stride (even)
  static constexpr int TILES = kBlocked ? NVEC * Runs::DOUBLES : kTile * TW;
  static constexpr int VT = DO_VAR ? kTile * (nb * p * p + 1) : 0;         // CTA tile: the nb p x p blocks per system + a zero
  static constexpr int THREADS = nb * 32;
  static constexpr int SMEM_DOUBLES = 2 * REC + TILES + VT + 2;
  static constexpr size_t smem_bytes = (size_t)SMEM_DOUBLES * sizeof(double);
  // resident CTAs per SM to aim for: as many as shared memory allows, registers capped at 128
  static constexpr int BY_SMEM = (int)((227 * 1024) / (smem_bytes + 1024));
  static constexpr int BY_REGS = 65536 / (THREADS * 128);
  static constexpr int MINB = BY_SMEM < BY_REGS ? (BY_SMEM < 1 ? 1 : BY_SMEM) : (BY_REGS < 1 ? 1 : BY_REGS);
};

// Per-warp work with the block index as a compile-time constant (so the block's Q, c, R are
// immediate constant-bank operands): Blk<A>::run is selected by a warp-uniform switch.
template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, int A>
struct BdwBlock {
  static constexpr int p = Fam::p, n = Fam::n, PS = Fam::PS;
  __device__ __forceinline__ static int step(const typename Fam::PriorT &P, bool terminal, const double *rec,
                                             const double *zt, double *mus, double *Ss, double *x) {
    const double *rmu = rec + A * p * kTile, *rS = rec + (n + A * PS) * kTile;
    const double *z = DO_SIM ? zt + A * p : nullptr;
    if (terminal) return reg::smooth_terminal<p, kTile, DO_MEAN, DO_VAR, DO_SIM>(rmu, rS, z, mus, Ss, x);
    return reg::smooth_step<p, kTile, DO_MEAN, DO_VAR, DO_SIM, true, Fam::kUT>(P.Q[A], P.c + A * p, P.R[A], rmu, rS, z,
                                                                                 mus, Ss, x);
  }
};

template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, int A = 0>
__device__ __forceinline__ int bdw_dispatch(int blk, const typename Fam::PriorT &P, bool terminal,
                                            const double *rec, const double *zt, double *mus, double *Ss,
                                            double *x) {
  if constexpr (A + 1 < Fam::nb) {
    if (blk != A) return bdw_dispatch<Fam, DO_MEAN, DO_VAR, DO_SIM, A + 1>(blk, P, terminal, rec, zt, mus, Ss, x);
  }
  return BdwBlock<Fam, DO_MEAN, DO_VAR, DO_SIM, A>::step(P, terminal, rec, zt, mus, Ss, x);
}

// Variance of one step.  Every warp parks its p x p block of each system in a CTA tile ([32][nb*p*p] doubles: only
// the non-zero entries); after the CTA barrier the warps share out the 32 systems and each writes WHOLE
// n*n-double runs (structural zeros synthesised, block entries read from the tile), lane k storing the
// 16-byte pieces k, k+32, ... of a run.  Whole runs matter: the memory system sustains long contiguous
// runs far better than a run assembled from nb pieces written at different times with ragged 8-byte
// edges (tools/microbench/write_pattern.cu).  For odd n a run starts on an odd double index for every
// other (system, step): pieces are cut on 16-byte boundaries of the OUTPUT ADDRESS, so the first / last
// piece of a run may be half (one 8-byte store).
template <class L>
struct BdwRunPlan {
  static constexpr int NN = L::n * L::n, NQ = (NN + 2) / 2, NIT = (NQ + 31) / 32;
  static constexpr int ZERO = L::nb * L::p * L::p;     // every tile row ends with a 0.0: structural zeros load it
  static constexpr int ROWW = ZERO + 1;
  // for address parity par and piece q = lane + 32*it: tile-row index of its two doubles, and what to
  // store (2 bits per (par, it): 0 nothing, 1 both doubles, 2 the lower one only, 3 the upper one only)
  int idx[2][NIT][2];
  unsigned kind;
  __device__ __forceinline__ void init(int lane) {
    constexpr int p = L::p, n = L::n;
    kind = 0;
#pragma unroll
    for (int par = 0; par < 2; par++)
#pragma unroll
      for (int it = 0; it < NIT; it++) {
        const int q = lane + 32 * it;
        bool ok[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int r = 2 * q - par + h;                         // double index inside the run: r = j*n + i
          ok[h] = q < NQ && r >= 0 && r < NN;
          int v = ZERO;
          if (ok[h]) {
            const int j = r / n, i = r - j * n;
            if (j / p == i / p) v = (j / p) * p * p + (j % p) * p + (i % p);
          }
          idx[par][it][h] = v;
        }
        const unsigned k = ok[0] && ok[1] ? 1u : ok[0] ? 2u : ok[1] ? 3u : 0u;
        kind |= k << (2 * (par * NIT + it));
      }
  }
};

template <class Fam, class L>
__device__ __forceinline__ void bdw_park_var(double *vt, const double *Ss, int blk, int lane) {
  constexpr int p = L::p;
  double *row = vt + lane * BdwRunPlan<L>::ROWW + blk * p * p;
#pragma unroll
  for (int jj = 0; jj < p; jj++)
#pragma unroll
    for (int ii = 0; ii < p; ii++) row[jj * p + ii] = Ss[sidx<p>(ii, jj)];
}

template <class L, int PAR>
__device__ __forceinline__ void bdw_emit_run(const double *src, const BdwRunPlan<L> &plan, double *run, int lane) {
  using Plan = BdwRunPlan<L>;
#pragma unroll
  for (int it = 0; it < Plan::NIT; it++) {
    const unsigned k = (plan.kind >> (2 * (PAR * Plan::NIT + it))) & 3u;
    const double v0 = src[plan.idx[PAR][it][0]], v1 = src[plan.idx[PAR][it][1]];
    double *dst = run - PAR + 2 * (lane + 32 * it);            // 16-byte aligned
    if (k == 1u) *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
    else if (k == 2u) dst[0] = v0;
    else if (k == 3u) dst[1] = v1;
  }
}

template <class Fam, class L>
__device__ __forceinline__ void bdw_emit_var(const double *vt, const BdwRunPlan<L> &plan, double *var_out, long b0,
                                             long B, long T1, int t, int blk, int lane) {
  using Plan = BdwRunPlan<L>;
  const int rows = (int)(B - b0 < kTile ? B - b0 : kTile);
  for (int s = blk; s < rows; s += L::nb) {
    double *run = var_out + ((b0 + s) * T1 + t) * (long)Plan::NN;
    const double *src = vt + s * Plan::ROWW;
    if ((reinterpret_cast<uintptr_t>(run) >> 3) & 1) bdw_emit_run<L, 1>(src, plan, run, lane);
    else bdw_emit_run<L, 0>(src, plan, run, lane);
  }
}

// Compact variance layouts (rodeo_cuda_set_var_layout): VL = 1 the n_meas diagonal (p, p) blocks -- exactly a row of
// the CTA tile -- VL = 2 the marginal variances.  The warps share out the systems as in bdw_emit_var and write whole
// runs of WV doubles (MSEIR: 45 or 15 instead of 225).
template <class L, int VL>
__device__ __forceinline__ void bdw_emit_var_compact(const double *vt, double *var_out, long b0, long B, long T1, int t,
                                                     int blk, int lane) {
  constexpr int p = L::p, nb = L::nb, WV = VL == 1 ? nb * p * p : nb * p;
  const int rows = (int)(B - b0 < kTile ? B - b0 : kTile);
  for (int s = blk; s < rows; s += nb) {
    double *run = var_out + ((b0 + s) * T1 + t) * (long)WV;
    const double *src = vt + s * BdwRunPlan<L>::ROWW;
    for (int e = lane; e < WV; e += 32) run[e] = VL == 1 ? src[e] : src[(e / p) * p * p + (e % p) * (p + 1)];
  }
}

template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, int VL = 0>
__global__ void __launch_bounds__(BdwLayout<Fam, DO_MEAN, DO_VAR, DO_SIM>::THREADS,
                                  BdwLayout<Fam, DO_MEAN, DO_VAR, DO_SIM>::MINB)
kalman_backward_bdw(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  using L = BdwLayout<Fam, DO_MEAN, DO_VAR, DO_SIM>;
  constexpr int p = L::p, n = L::n, PS = L::PS, REC = L::REC;
  using Runs = typename L::Runs;
  extern __shared__ __align__(128) double smem[];
  double *ring = smem;
  double *tiles = ring + 2 * REC;                       // mean / draw rows: K-step blocks (BlockedRuns)
  double *vtile = tiles + L::TILES;                     // the systems' diagonal variance blocks (CTA-wide)
  uint64_t *bar = reinterpret_cast<uint64_t *>(vtile + L::VT);
  const int tid = threadIdx.x, blk = tid / 32, lane = tid % 32;
  const long tile = blockIdx.x;
  const int N = P.n_eval;   // checkpoint interval 1: record r holds time index N - r, R = N records
  const long T1 = (long)N + 1, b0 = tile * kTile, b = b0 + lane;
  const bool valid = b < a.B;
  const double *hist_tile = a.hist + (size_t)tile * N * REC;
  constexpr uint32_t kRecBytes = REC * sizeof(double);

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 2; k++)
      if (k < N) {
        mbar_arrive_expect_tx(&bar[k], kRecBytes);
        bulk_g2s(ring + k * REC, hist_tile + (size_t)k * REC, kRecBytes, &bar[k]);
      }
  }
  double mus[p], Ss[PS], x[p];
  int fail = 0;
  Runs rm, rx;
  if (L::kBlocked && DO_MEAN) rm.init(tiles, a.mu_out, b0, a.B, T1, tid);
  if (L::kBlocked && DO_SIM) rx.init(tiles + (DO_MEAN ? Runs::DOUBLES : 0), a.x_out, b0, a.B, T1, tid);
  BdwRunPlan<L> vplan;
  if (DO_VAR) {
    vplan.init(lane);
    if (blk == 0) vtile[lane * BdwRunPlan<L>::ROWW + BdwRunPlan<L>::ZERO] = 0.0;   // ordered by the first CTA barrier
  }
  const double *zb = DO_SIM ? a.z + (valid ? b : b0) * a.z_stride : nullptr;
  uint32_t phase = 0;
  for (int t = N; t >= 0; t--) {
    if (t >= 1) {
      const int k = (N - t) & 1;
      mbar_wait(&bar[k], (phase >> k) & 1u);
      phase ^= 1u << k;
      const double *rec = ring + k * REC + lane;
      const double *zt = DO_SIM ? zb + (long)(t - 1) * n : nullptr;
      const int f = bdw_dispatch<Fam, DO_MEAN, DO_VAR, DO_SIM>(blk, P, t == N, rec, zt, mus, Ss, x);
      if (f && !fail) fail = blk * p + f;
      fence_proxy_async();  // this thread's reads of slot k are complete before the TMA may overwrite it
      __syncthreads();      // every warp is done with ring slot k (and with last step's tiles)
      if (tid == 0 && N - t + 2 < N) {  // refill the slot with record r + 2 (time index t - 2)
        mbar_arrive_expect_tx(&bar[k], kRecBytes);
        bulk_g2s(ring + k * REC, hist_tile + (size_t)(N - t + 2) * REC, kRecBytes, &bar[k]);
      }
    } else {  // time index 0: the initial state is known, Sigma_0 = 0 (KalmanODE.pyx:445, 409)
#pragma unroll
      for (int i = 0; i < p; i++) {
        const double v = valid ? a.x0[b * n + blk * p + i] : 0.0;
        if (DO_MEAN) mus[i] = v;
        if (DO_SIM) x[i] = v;
      }
      if (DO_VAR) {
#pragma unroll
        for (int i = 0; i < PS; i++) Ss[i] = 0.0;
      }
      __syncthreads();
    }
    // ---- park this block's part of every output in the CTA tiles, one barrier, then store ----------
    if constexpr (L::kBlocked) {
      if (DO_MEAN) rm.template put_part<p>(lane, t, blk * p, mus);
      if (DO_SIM) rx.template put_part<p>(lane, t, blk * p, x);
      bdw_park_var<Fam, L>(vtile, Ss, blk, lane);
      __syncthreads();
      if (DO_MEAN) rm.store(t, tid);      // the 4-step blocks that became complete at this step
      if (DO_SIM) rx.store(t, tid);
      if constexpr (VL == 0) bdw_emit_var<Fam, L>(vtile, vplan, a.var_out, b0, a.B, T1, t, blk, lane);
      else bdw_emit_var_compact<L, VL>(vtile, a.var_out, b0, a.B, T1, t, blk, lane);
    } else {
      constexpr int TW = L::TW, NT = L::THREADS;
      double *row = tiles + lane * TW;
      if (DO_MEAN) {
#pragma unroll
        for (int i = 0; i < p; i++) row[blk * p + i] = mus[i];
      }
      if (DO_SIM) {
#pragma unroll
        for (int i = 0; i < p; i++) row[(DO_MEAN ? n : 0) + blk * p + i] = x[i];
      }
      __syncthreads();
      if (DO_MEAN) flush_part<n, 0, n, 0, TW, NT>(tiles, a.mu_out, b0, a.B, T1, t, tid);
      if (DO_SIM) flush_part<n, 0, n, (DO_MEAN ? n : 0), TW, NT>(tiles, a.x_out, b0, a.B, T1, t, tid);
    }
  }
  if (valid && a.status && fail) atomicCAS(a.status + b, 0, fail);
}

}  // namespace rodeo

namespace rodeo {

template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, int VL = 0>
cudaError_t launch_backward_bdw(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  using L = BdwLayout<Fam, DO_MEAN, DO_VAR, DO_SIM>;
  auto kern = kalman_backward_bdw<Fam, DO_MEAN, DO_VAR, DO_SIM, VL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem_bytes);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((a.B + kTile - 1) / kTile);
  kern<<<grid, L::THREADS, L::smem_bytes, s>>>(make_prior<Fam>(h), a);
  return cudaGetLastError();
}

}  // namespace rodeo
