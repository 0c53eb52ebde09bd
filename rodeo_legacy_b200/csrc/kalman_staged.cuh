// kalman_staged.cuh -- backward (smoother) kernel of the register family with the
// filter history staged through shared memory by bulk async copies (TMA, cp.async.bulk +
// mbarrier) and the outputs transposed through shared memory into coalesced 16-byte stores.
//
// One warp owns one tile of 32 systems (one thread per system) for the whole backward pass:
//   * its history record for one (checkpointed) time index is NE*32 contiguous doubles in HBM
//     (layout in kalman_common.cuh).  Lane 0 issues ONE cp.async.bulk per record into a two-slot
//     ring in shared memory, two records ahead of the one being processed, so the
//     DRAM latency of the read stream is hidden without spending registers on prefetch;
//     each lane then reads its own column of the record (bank-conflict free).
//   * every lane writes the step's outputs (mean, covariance, draw) into its own row of a
//     staging tile; the warp then stores the tile so that each system's contiguous run in
//     the user-visible layout  out[b][t][...]  is written by neighbouring lanes as 16-byte
//     pieces -- full 32-byte sectors instead of 32 scattered 8-byte stores per instruction.
//   * with checkpointing (CK > 1) the filter is re-run from each checkpoint record into CK-1
//     scratch records in shared memory (same layout, each lane touches only its own column),
//     which the smoother then consumes exactly like loaded records.
// Warps never synchronise with each other (no __syncthreads after the prologue).
#pragma once
#include <cstdint>

#include "kalman_tps.cuh"
#include "registry.cuh"

namespace rodeo {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA, 1-D), completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kStagedWarps = 2;  // warps per CTA
constexpr int kStagedThreads = kStagedWarps * 32;

template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
struct StagedLayout {
  static constexpr int n = Fam::n, NE = Fam::NE;
  static constexpr int REC = NE * kTile;                                         // doubles per record
  static constexpr int OUTW = (DO_MEAN ? n : 0) + (DO_VAR ? n * n : 0) + (DO_SIM ? n : 0);
  static constexpr int ROW = (OUTW + 1) / 2 * 2;                                 // row stride (even)
  // per warp: 2 ring slots (checkpoint records, TMA), CK-1 scratch records (re-run filter),
  // the output staging tile, 2 mbarriers
  static constexpr int WARP_DOUBLES = (2 + (CK - 1)) * REC + kTile * ROW + 2;
  static constexpr size_t smem_bytes = (size_t)kStagedWarps * WARP_DOUBLES * sizeof(double);
  // No min-blocks launch bound: capping registers (tried: 168 / 128 per thread for the
  // block-diagonal family) makes ptxas spill inside the unrolled step and DOUBLED the kernel
  // time on B200 (2.52 -> 5.12 ms, FN solve_mv); 254 registers and 8 warps / SM is faster.
  static constexpr int MINB = 1;
};

// Store one output array from the staging tile: every system s of the tile owns a run of
// W doubles at stage[s*ROW + OFF ..]; its destination is out + ((b0+s)*T1 + t)*W.
template <int W, int OFF, int ROW>
__device__ __forceinline__ void flush_runs(const double *stage, double *out, long b0, long B, long T1, int t,
                                           int lane) {
  constexpr bool V2 = (W % 2 == 0) && (OFF % 2 == 0);
  if (V2) {
    constexpr int C = W / 2;
#pragma unroll
    for (int q0 = 0; q0 < kTile * C; q0 += kTile) {
      const int q = q0 + lane;
      if (q < kTile * C) {
        const int s = q / C, c = q - s * C;
        if (b0 + s < B) {
          const double2 v = *reinterpret_cast<const double2 *>(stage + s * ROW + OFF + 2 * c);
          *reinterpret_cast<double2 *>(out + ((b0 + s) * T1 + t) * W + 2 * c) = v;
        }
      }
    }
  } else {
#pragma unroll
    for (int q0 = 0; q0 < kTile * W; q0 += kTile) {
      const int q = q0 + lane;
      if (q < kTile * W) {
        const int s = q / W, c = q - s * W;
        if (b0 + s < B) out[((b0 + s) * T1 + t) * W + c] = stage[s * ROW + OFF + c];
      }
    }
  }
}

// Emit time index t of the whole tile: every lane writes its row, the warp stores the tile.
template <class Fam, class L, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
__device__ __forceinline__ void emit_staged(double *stage, const SolveArgs &a, long b0, long T1, int t, int lane,
                                            const double *mus, const double *Ss, const double *x) {
  constexpr int n = Fam::n, ROW = L::ROW;
  constexpr int OFF_VAR = DO_MEAN ? n : 0, OFF_X = OFF_VAR + (DO_VAR ? n * n : 0);
  double *row = stage + lane * ROW;
  if (DO_MEAN) {
#pragma unroll
    for (int i = 0; i < n; i++) row[i] = mus[i];
  }
  if (DO_VAR) {
#pragma unroll
    for (int j = 0; j < n; j++)
#pragma unroll
      for (int i = 0; i < n; i++) row[OFF_VAR + j * n + i] = Fam::var_at(Ss, i, j);
  }
  if (DO_SIM) {
#pragma unroll
    for (int i = 0; i < n; i++) row[OFF_X + i] = x[i];
  }
  __syncwarp();
  if (DO_MEAN) flush_runs<n, 0, ROW>(stage, a.mu_out, b0, a.B, T1, t, lane);
  if (DO_VAR) flush_runs<n * n, OFF_VAR, ROW>(stage, a.var_out, b0, a.B, T1, t, lane);
  if (DO_SIM) flush_runs<n, OFF_X, ROW>(stage, a.x_out, b0, a.B, T1, t, lane);
  __syncwarp();
}

template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
__global__ void __launch_bounds__(kStagedThreads, StagedLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>::MINB)
kalman_backward_staged(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  using L = StagedLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>;
  using Ode = typename Fam::Ode;
  constexpr int n = Fam::n, NS = Fam::NS, NE = Fam::NE, REC = L::REC;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long tile = (long)blockIdx.x * kStagedWarps + warp;
  const long n_tiles = (a.B + kTile - 1) / kTile;
  if (tile >= n_tiles) return;  // warp-uniform; warps are independent from here on
  double *ring = smem + (size_t)warp * L::WARP_DOUBLES;
  double *scratch = ring + 2 * REC;
  double *stage = scratch + (CK - 1) * REC;
  uint64_t *bar = reinterpret_cast<uint64_t *>(stage + kTile * L::ROW);
  const int N = P.n_eval, R = n_records(N, CK);
  const long T1 = (long)N + 1, b0 = tile * kTile, b = b0 + lane;
  const bool valid = b < a.B;
  const long bs = valid ? b : b0;  // a safe system index for loads of inactive lanes
  const double *hist_tile = a.hist + (size_t)tile * R * REC;
  constexpr uint32_t kRecBytes = REC * sizeof(double);

  if (lane == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 2; k++)
      if (k < R) {
        mbar_arrive_expect_tx(&bar[k], kRecBytes);
        bulk_g2s(ring + k * REC, hist_tile + (size_t)k * REC, kRecBytes, &bar[k]);
      }
  }
  double mus[n], Ss[NS], x[n], th[NT];
  int fail = 0;
  const double *zb = DO_SIM ? a.z + bs * a.z_stride : nullptr;
  if (CK > 1) {
#pragma unroll
    for (int i = 0; i < NT; i++) th[i] = (Ode::n_theta > 0) ? a.theta[bs * Ode::n_theta + i] : 0.0;
  }
  uint32_t phase = 0;
  // wait for record r (ring slot r & 1); returns this lane's column of it
  auto acquire = [&](int r) -> const double * {
    const int k = r & 1;
    mbar_wait(&bar[k], (phase >> k) & 1u);
    phase ^= 1u << k;
    return ring + k * REC + lane;
  };
  // every lane is done with record r: refill its slot with record r + 2
  auto release = [&](int r) {
    __syncwarp();
    if (lane == 0 && r + 2 < R) {
      const int k = r & 1;
      mbar_arrive_expect_tx(&bar[k], kRecBytes);
      bulk_g2s(ring + k * REC, hist_tile + (size_t)(r + 2) * REC, kRecBytes, &bar[k]);
    }
  };
  {  // terminal time index N = record 0
    const double *rec = acquire(0);
    fail = Fam::template bw_terminal<kTile, DO_MEAN, DO_VAR, DO_SIM>(P, rec, DO_SIM ? zb + (long)(N - 1) * n : nullptr,
                                                                     mus, Ss, x);
    release(0);
    emit_staged<Fam, L, DO_MEAN, DO_VAR, DO_SIM>(stage, a, b0, T1, N, lane, mus, Ss, x);
  }
  int r = 1;
  for (int t_hi = N; t_hi > 1;) {  // this group smooths time indices t_hi-1 .. max(t_hi-CK, 1)
    const int base = t_hi - CK;
    if (CK == 1) {
      const double *rec = acquire(r);
      const int f = Fam::template bw_step<kTile, DO_MEAN, DO_VAR, DO_SIM>(
          P, rec, DO_SIM ? zb + (long)(base - 1) * n : nullptr, mus, Ss, x);
      if (f && !fail) fail = f;
      release(r);
      emit_staged<Fam, L, DO_MEAN, DO_VAR, DO_SIM>(stage, a, b0, T1, base, lane, mus, Ss, x);
    } else {
      const int t0 = base >= 1 ? base : 0;
      const int nfwd = t_hi - 1 - t0;  // re-run time indices t0+1 .. t_hi-1 into the scratch records
      const double *rec = nullptr;
      {
        double fm[n], fS[NS];
        if (base >= 1) {
          rec = acquire(r);
#pragma unroll
          for (int i = 0; i < n; i++) fm[i] = rec[i * kTile];
#pragma unroll
          for (int i = 0; i < NS; i++) fS[i] = rec[(n + i) * kTile];
        } else {
#pragma unroll
          for (int i = 0; i < n; i++) fm[i] = a.x0[bs * n + i];
#pragma unroll
          for (int i = 0; i < NS; i++) fS[i] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < CK - 1; j++) {
          if (j < nfwd) {
            Fam::template step<IG_PROBDE>(P, t0 + j, th, nullptr, fm, fS);
            double *sr = scratch + j * REC + lane;
#pragma unroll
            for (int i = 0; i < n; i++) sr[i * kTile] = fm[i];
#pragma unroll
            for (int i = 0; i < NS; i++) sr[(n + i) * kTile] = fS[i];
          }
        }
      }
#pragma unroll
      for (int j = CK - 2; j >= 0; j--) {
        if (j < nfwd) {
          const int t = t0 + j + 1;
          const int f = Fam::template bw_step<kTile, DO_MEAN, DO_VAR, DO_SIM>(
              P, scratch + j * REC + lane, DO_SIM ? zb + (long)(t - 1) * n : nullptr, mus, Ss, x);
          if (f && !fail) fail = f;
          emit_staged<Fam, L, DO_MEAN, DO_VAR, DO_SIM>(stage, a, b0, T1, t, lane, mus, Ss, x);
        }
      }
      if (base >= 1) {
        const int f = Fam::template bw_step<kTile, DO_MEAN, DO_VAR, DO_SIM>(
            P, rec, DO_SIM ? zb + (long)(base - 1) * n : nullptr, mus, Ss, x);
        if (f && !fail) fail = f;
        release(r);
        emit_staged<Fam, L, DO_MEAN, DO_VAR, DO_SIM>(stage, a, b0, T1, base, lane, mus, Ss, x);
      }
    }
    t_hi = base >= 1 ? base : 1;
    r++;
  }
  {  // time index 0: the initial state is known, Sigma_0 = 0 (KalmanODE.pyx:445, 409)
#pragma unroll
    for (int i = 0; i < n; i++) {
      const double v = valid ? a.x0[b * n + i] : 0.0;
      mus[i] = v;
      x[i] = v;
    }
#pragma unroll
    for (int i = 0; i < NS; i++) Ss[i] = 0.0;
    emit_staged<Fam, L, DO_MEAN, DO_VAR, DO_SIM>(stage, a, b0, T1, 0, lane, mus, Ss, x);
  }
  if (valid && a.status && fail && a.status[b] == 0) a.status[b] = fail;
}

template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
cudaError_t launch_backward_staged(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  using L = StagedLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>;
  auto kern = kalman_backward_staged<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem_bytes);
  if (e != cudaSuccess) return e;
  const long n_tiles = (a.B + kTile - 1) / kTile;
  const unsigned grid = (unsigned)((n_tiles + kStagedWarps - 1) / kStagedWarps);
  kern<<<grid, kStagedThreads, L::smem_bytes, s>>>(make_prior<Fam>(h), a);
  return cudaGetLastError();
}

}  // namespace rodeo
