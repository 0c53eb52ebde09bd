// kalman_tpx.cuh -- draw-only backward kernel (solve_sim) of the block-diagonal family with >= 3 blocks:
// one THREAD PER DIAGONAL BLOCK, checkpointed history.  Written for BASELINE config 3 (Lorenz63, 16,384 x0,
// n_eval 5000, solve_sim), the config furthest from its bytes in round 1:
//   * 16,384 systems x 3 blocks are 49,152 independent 5,000-step chains = 10.4 warps per SM: the step is bound by
//     the dependent FP64 chain of one block step (two 3 x 3 Cholesky factorisations in sequence: 6 x 67 cycles of
//     rsqrt alone, tools/microbench/fp64_latency.cu), not by throughput or bandwidth (ncu: DRAM 40 %).  Everything
//     else is therefore kept OFF that chain:
//   * the warp-per-block kernel (kalman_bdw.cuh) needs the full history because re-running the filter means a
//     CTA barrier per step; here the blocks of a system sit in one warp and exchange the m predicted values the
//     right-hand side needs with shuffles, so checkpoint interval 2 works: half the history traffic
//     (Lorenz63: 17.7 -> 8.9 GB written and read);
//   * history: each lane copies its block's next checkpoint record into its own shared-memory slots with cp.async
//     (LDGSTS) one group ahead and picks it up with cp.async.wait_group + LDS.  With register prefetches (plain
//     loads a group ahead) the in-flight DRAM loads shared a scoreboard with the z loads of the current step: the
//     DFMA that consumes z waited for the PREFETCH to land (ncu: 27 % of all stall samples on that one instruction,
//     unchanged when the z loads were hoisted).  5.25 -> 4.79 ms, DRAM reads 17.9 -> 9.2 GB;
//   * draws: every step the lanes park their block's p values in a per-warp shared-memory image [system][n]
//     (double buffered: one warp barrier per step) and the warp writes it back with plain coalesced stores,
//     consecutive lanes = consecutive doubles of a system's 8n-byte run.  The first version kept per-system FIFOs
//     and issued one cp.async.bulk per completed 256-byte granule (what the HBM-bound headline smoother needs):
//     1.2 G more instructions (serialised UBLKCP issue loops, proxy fences, wait_group), 5.74 vs 5.24 ms;
//   * z[:, t-1] is loaded with volatile loads at the top of the step; right-hand sides with more than three
//     parameters re-read theta from per-lane shared-memory slots (registers: the cap below).
// Lane = blk * SPW + s, SPW = 32 / n_blocks systems per warp and slot; warp w owns systems
// [w * SPW * U, (w + 1) * SPW * U), system of (slot u, lane s) = w * SPW * U + u * SPW + s (U = 1 is what is built:
// two systems per lane for instruction-level parallelism measured 11.3 vs 5.3 ms, ptxas serialises the streams at
// 234 registers).  History in the plain (unpermuted) tile order.  Warps never synchronise with each other.
#pragma once
#include <cstdint>

#include "kalman_tpb.cuh"

namespace rodeo {

template <class Fam, int CK, int U>
struct TpxLayout {
  static constexpr int p = Fam::p, nb = Fam::nb, n = Fam::n, PS = Fam::PS;
  static constexpr int SPW = 32 / nb, LANES = SPW * nb;
  static constexpr int WARPS = 2, THREADS = WARPS * 32;
  // U = 1: 12 resident warps per SM (<= 170 registers) hold a whole 16,384-system Lorenz63 batch (11.1 warps per
  // SM) in ONE wave; at the 190 registers ptxas would take, 10 fit and the launch needs two
  static constexpr int MINB = U == 1 ? 6 : 1;
  static constexpr int IMG = SPW * n;                       // doubles of one staging image of the draws
  static constexpr int NST = (IMG + 31) / 32;               // store instructions per step
  static constexpr int REC = p + PS;                        // doubles of one block's history record
  static constexpr int NTH = Fam::Ode::n_theta;
  // six parameters (MSEIR) in registers push the five-block instance over the 168-register cap (44 bytes of
  // spills, solve_sim smoother 3.61 ms; from shared memory 3.14 ms)
  static constexpr bool kThetaSmem = CK > 1 && NTH > 3;
  // dynamic shared memory: [draw images: WARPS x U x 2 x IMG][record slots: U x 2 stages x REC x THREADS][theta]
  static constexpr size_t stage_off = ((size_t)WARPS * U * 2 * IMG * 8 + 15) / 16 * 16;
  static constexpr size_t theta_off = stage_off + (size_t)THREADS * U * 2 * REC * 8;
  static constexpr size_t smem_bytes = theta_off + (kThetaSmem ? (size_t)THREADS * U * NTH * 8 : 0);
  static_assert(CK == 1 || CK == 2, "thread-per-block smoother: checkpoint interval 1 or 2");
};

template <class Fam, int CK, int U, bool UNIFORM>
__global__ void __launch_bounds__(TpxLayout<Fam, CK, U>::THREADS, TpxLayout<Fam, CK, U>::MINB)
kalman_backward_tpx(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  using L = TpxLayout<Fam, CK, U>;
  using Ode = typename Fam::Ode;
  constexpr int p = L::p, nb = L::nb, n = L::n, PS = L::PS, NE = Fam::NE, SPW = L::SPW;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long wg = (long)blockIdx.x * L::WARPS + warp;
  const long b_first = wg * SPW * U;
  if (b_first >= a.B) return;                              // warp-uniform; warps are independent
  const bool lane_on = lane < L::LANES;
  const int blk = lane_on ? lane / SPW : 0, s = lane_on ? lane % SPW : 0;
  const int N = P.n_eval, R = n_records(N, CK);
  const long T1 = (long)N + 1;
  const BlockPrior<Fam, UNIFORM> bp(P, blk);

  long b[U], bs[U];
  bool valid[U];
  const double *hb[U], *zb[U];
  double th[U][NT], x0b[U][p];
#pragma unroll
  for (int u = 0; u < U; u++) {
    b[u] = b_first + u * SPW + s;
    valid[u] = lane_on && b[u] < a.B;
    bs[u] = b[u] < a.B ? b[u] : b_first;                   // a safe system for the loads of idle lanes
    hb[u] = a.hist + hist_offset(bs[u], 0, R, NE);
    zb[u] = a.z + bs[u] * a.z_stride + blk * p;
#pragma unroll
    for (int e = 0; e < p; e++) x0b[u][e] = ld_once(a.x0 + bs[u] * n + blk * p + e);
#pragma unroll
    for (int e = 0; e < NT; e++) {
      th[u][e] = (Ode::n_theta > 0 && CK > 1) ? ld_once(a.theta + bs[u] * Ode::n_theta + e) : 0.0;
      if constexpr (L::kThetaSmem)
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(smem_u32(smem_raw) + (uint32_t)L::theta_off +
                                                     (uint32_t)(((u * NT + e) * L::THREADS + (int)threadIdx.x) * 8)),
                     "d"(th[u][e])
                     : "memory");
    }
  }
  auto load_rec = [&](int u, int r, double *mu, double *S) {
    const double *h = hb[u] + (size_t)r * NE * kTile;
#pragma unroll
    for (int e = 0; e < p; e++) mu[e] = ld_stream(h + (blk * p + e) * kTile);
#pragma unroll
    for (int e = 0; e < PS; e++) S[e] = ld_stream(h + (n + blk * PS + e) * kTile);
  };
  // cp.async prefetch of record r into stage r & 1 / pick-up from it; slot of element e: [u][stage][e][thread]
  const uint32_t stage_base = smem_u32(smem_raw) + (uint32_t)L::stage_off + (uint32_t)threadIdx.x * 8u;
  auto stage_addr = [&](int u, int st, int e) {
    return stage_base + (uint32_t)((((u * 2 + st) * L::REC + e) * L::THREADS) * 8);
  };
  auto prefetch_rec = [&](int u, int r) {
    const double *h = hb[u] + (size_t)r * NE * kTile;
    const int st = r & 1;
#pragma unroll
    for (int e = 0; e < p; e++)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stage_addr(u, st, e)), "l"(h + (blk * p + e) * kTile) : "memory");
#pragma unroll
    for (int e = 0; e < PS; e++)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stage_addr(u, st, p + e)), "l"(h + (n + blk * PS + e) * kTile)
                   : "memory");
  };
  auto fetch_rec = [&](int u, int r, double *mu, double *S) {
    const int st = r & 1;
#pragma unroll
    for (int e = 0; e < p; e++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(mu[e]) : "r"(stage_addr(u, st, e)) : "memory");
#pragma unroll
    for (int e = 0; e < PS; e++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(S[e]) : "r"(stage_addr(u, st, p + e)) : "memory");
  };
  // staging images of the draws (shared-state addresses) and this lane's store offsets: double k * 32 + lane of an image =
  // element el of system sy of the warp, stored at xw[(sy * T1 + t) * n + el]
  const uint32_t img_base = smem_u32(smem_raw) + (uint32_t)(warp * U) * 2u * L::IMG * 8u;
  const uint32_t img_wr = img_base + (uint32_t)(s * n + blk * p) * 8u, img_rd = img_base + (uint32_t)lane * 8u;
  uint32_t img_par = 0;                                     // byte offset of the image in use (0 or IMG * 8)
  double *const xw = a.x_out + b_first * T1 * n;            // warp-uniform
  int ooff[U][L::NST];
  bool ov[U][L::NST];
#pragma unroll
  for (int u = 0; u < U; u++)
#pragma unroll
    for (int k = 0; k < L::NST; k++) {
      const int idx = k * 32 + lane, sy = idx / n, el = idx - sy * n;
      ov[u][k] = idx < L::IMG && b_first + u * SPW + sy < a.B;
      ooff[u][k] = (int)((u * SPW + sy) * T1 * n) + el;
    }
  // park the draws of time index t in the image of this step's parity, one warp barrier, coalesced stores
  // (double buffered: the reads of the other image precede this step's barrier)
  auto emit = [&](int t, double (*x)[p]) {
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int e = 0; e < p; e++)
        if (lane_on)
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(img_wr + img_par + (uint32_t)((u * 2 * L::IMG + e) * 8)), "d"(x[u][e])
                       : "memory");
    __syncwarp();
    const int tn = t * n;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int k = 0; k < L::NST; k++) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(img_rd + img_par + (uint32_t)((u * 2 * L::IMG + k * 32) * 8)) : "memory");
        if (ov[u][k]) xw[ooff[u][k] + tn] = v;
      }
    img_par ^= (uint32_t)(L::IMG * 8);
  };
  // z[:, t-1], this block's rows.  Volatile loads: issued HERE, at the top of the step; left to the compiler they sink
  // to their use at the end of the step (x = m~ + L z) and their L2 latency lands on the chain -- ncu: 27 % of
  // all stall samples on that one DFMA
  auto z_at = [&](int u, int t, double *zt) {
#pragma unroll
    for (int e = 0; e < p; e++) zt[e] = ld_once(zb[u] + (long)(t - 1) * n + e);
  };

  double x[U][p], cm[U][p], cS[U][PS], mus_unused[p], Ss_unused[PS];
  int fail[U];
#pragma unroll
  for (int u = 0; u < U; u++) fail[u] = 0;
#pragma unroll
  for (int u = 0; u < U; u++) {
    load_rec(u, 0, cm[u], cS[u]);
    if (R > 1) prefetch_rec(u, 1);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
  for (int u = 0; u < U; u++) {   // terminal time index N = record 0
    double zt[p];
    z_at(u, N, zt);
    const int f = reg::smooth_terminal<p, 1, false, false, true>(cm[u], cS[u], zt, mus_unused, Ss_unused, x[u]);
    if (f && !fail[u]) fail[u] = f;
  }
  emit(N, x);
  int r = 1;
  for (int t_hi = N; t_hi > 1;) {   // this group draws time indices t_hi-1 .. max(t_hi-CK, 1)
    const int base = t_hi - CK;
    // z of the group's FIRST time index is issued before the asynchronous copies of this group: loaded after them,
    // its consumer waited for them (12 % of the stall samples; 4.31 -> 4.11 ms).  Hoisting the second index as
    // well costs registers (spills at the 168 cap: 4.20 ms; with theta moved to shared memory instead: 4.25 ms).
    double ztA[U][p];
    if (CK == 2) {
#pragma unroll
      for (int u = 0; u < U; u++) z_at(u, (base >= 1 ? base : 0) + 1, ztA[u]);
    }
    if (base >= 1) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");            // record r (copied a group ago) is in its stage
#pragma unroll
      for (int u = 0; u < U; u++) {
        fetch_rec(u, r, cm[u], cS[u]);
        if (r + 1 < R) prefetch_rec(u, r + 1);                        // one group ahead, into the other stage
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (CK == 2) {
      const int t0 = base >= 1 ? base : 0;
      if (base < 1) {   // the known initial state (KalmanODE.pyx:293-297)
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
          for (int e = 0; e < p; e++) cm[u][e] = x0b[u][e];
#pragma unroll
          for (int e = 0; e < PS; e++) cS[u][e] = 0.0;
        }
      }
      if (t_hi - 1 > t0) {
        // re-run the filter t0 -> t0 + 1 (BdFam::step, this block's share; the right-hand side needs every
        // block's predicted first component: one shuffle per block), then draw time index t0 + 1
        const int t = t0 + 1;
        const double tt = P.tmin + (P.tmax - P.tmin) * (t0 + 1) / P.n_eval;
#pragma unroll
        for (int u = 0; u < U; u++) {
          double fm[p], fS[PS];
          const double *zt = ztA[u];
          reg::predict_mean<p, Fam::kUT>(bp.Q(), bp.c(), cm[u], fm);
          reg::predict_var<p, false, Fam::kUT>(bp.Q(), bp.R(), cS[u], fS, nullptr);
          double X[n], y[Fam::m];
#pragma unroll
          for (int e = 0; e < n; e++) X[e] = 0.0;
#pragma unroll
          for (int k = 0; k < nb; k++) X[k * p] = __shfl_sync(kFull, fm[0], k * SPW + s);
          if constexpr (L::kThetaSmem) {   // this lane's own slots: no barrier needed
            double thl[NT];
#pragma unroll
            for (int e = 0; e < NT; e++)
              asm volatile("ld.shared.f64 %0, [%1];"
                           : "=d"(thl[e])
                           : "r"(smem_u32(smem_raw) + (uint32_t)L::theta_off +
                                 (uint32_t)(((u * NT + e) * L::THREADS + (int)threadIdx.x) * 8))
                           : "memory");
            Ode::template eval<n>(X, tt, thl, y);
          } else {
            Ode::template eval<n>(X, tt, th[u], y);
          }
          double ya = y[0];
#pragma unroll
          for (int k = 1; k < nb; k++) ya = (blk == k) ? y[k] : ya;
          reg::update_block<p, Fam::kWK>(bp.w(), fm, fS, ya, false);
          const int f = reg::smooth_step<p, 1, false, false, true, false, Fam::kUT>(
              bp.Q(), bp.c(), bp.R(), fm, fS, zt, mus_unused, Ss_unused, x[u]);
          if (f && !fail[u]) fail[u] = f;
        }
        emit(t, x);
      }
    }
    if (base >= 1) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        double zt[p];
        z_at(u, base, zt);
        const int f = reg::smooth_step<p, 1, false, false, true, false, Fam::kUT>(
            bp.Q(), bp.c(), bp.R(), cm[u], cS[u], zt, mus_unused, Ss_unused, x[u]);
        if (f && !fail[u]) fail[u] = f;
      }
      emit(base, x);
    }
    t_hi = base >= 1 ? base : 1;
    r++;
  }
#pragma unroll
  for (int u = 0; u < U; u++)   // time index 0: the initial state is known (KalmanODE.pyx:409)
#pragma unroll
    for (int e = 0; e < p; e++) x[u][e] = x0b[u][e];
  emit(0, x);
#pragma unroll
  for (int u = 0; u < U; u++)
    if (valid[u] && a.status && fail[u]) atomicCAS(a.status + b[u], 0, blk * p + fail[u]);
}

// A thread-per-block FORWARD filter for >= 3 blocks was measured and dropped: with 10 systems per warp its history
// stores are 80-byte pieces (partial 32-byte sectors: Lorenz63 3.9 ms vs 1.9 ms for the thread-per-system kernel,
// MSEIR 7.6 vs 2.0 ms), with 8 systems per warp (aligned 64-byte pieces) still 2.08 vs 1.88 ms -- the forward pass is
// bound by how its history rows reach DRAM, and one thread per system writes whole 256-byte rows per instruction.

template <class Fam, int CK, int U>
cudaError_t launch_backward_tpx(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  using L = TpxLayout<Fam, CK, U>;
  if (a.perm16) return cudaErrorInvalidValue;
  const typename Fam::PriorT P = make_prior<Fam>(h);
  const bool uni = blocks_uniform<Fam>(P);
  auto kern = uni ? kalman_backward_tpx<Fam, CK, U, true> : kalman_backward_tpx<Fam, CK, U, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem_bytes);
  if (e != cudaSuccess) return e;
  const long per_warp = (long)L::SPW * U;
  const long n_warps = (a.B + per_warp - 1) / per_warp;
  const unsigned grid = (unsigned)((n_warps + L::WARPS - 1) / L::WARPS);
  kern<<<grid, L::THREADS, L::smem_bytes, s>>>(P, a);
  return cudaGetLastError();
}

}  // namespace rodeo
