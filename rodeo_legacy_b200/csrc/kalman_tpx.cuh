// kalman_tpx.cuh -- draw-only backward kernel (solve_sim) of the block-diagonal family with >= 3 blocks:
// one THREAD PER DIAGONAL BLOCK, U systems per lane, checkpointed history.  Written for BASELINE config 3
// (Lorenz63, 16,384 x0, n_eval 5000, solve_sim), the config furthest from its bytes in round 1:
//   * 16,384 systems x 3 blocks are 49,152 independent 5,000-step chains = 10.4 warps per SM: the step is bound by
//     the LATENCY of one block step (two 3 x 3 Cholesky factorisations in sequence), not by throughput.  Each
//     lane therefore runs the blocks of U = 2 different systems interleaved (independent instruction streams:
//     the compiler overlaps their dependency chains), which halves the latency per system.
//   * the warp-per-block kernel (kalman_bdw.cuh) needs the full history because re-running the filter means a
//     CTA barrier per step; here the blocks of a system sit in one warp and exchange the m predicted values the
//     right-hand side needs with shuffles, so checkpoint interval 2 works: half the history traffic
//     (Lorenz63: 17.7 -> 8.9 GB written and read).
//   * draws leave through a per-system FIFO in shared memory as whole 256-byte-aligned granules, one
//     cp.async.bulk per granule issued by the lane of block 0 (72-byte runs: a granule every 3.6 steps per
//     system, so the per-lane issue loop that was too slow for 288-byte variance runs is cheap here).
// Lane = blk * SPW + s, SPW = 32 / n_blocks systems per warp and slot; warp w owns systems
// [w * SPW * U, (w + 1) * SPW * U), system of (slot u, lane s) = w * SPW * U + u * SPW + s.  History in the plain
// (unpermuted) tile order.  Warps never synchronise with each other.
#pragma once
#include <cstdint>

#include "kalman_tpb.cuh"

namespace rodeo {

__device__ __forceinline__ void bulk_s2g(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem),
               "r"(bytes)
               : "memory");
}

// FIFO of one output stream (W doubles per system and step) of ONE system: the image of the not-yet-stored
// bytes of its trajectory, addressed by global byte offset modulo the FIFO size.
// EK = steps between emissions: the FIFO holds what an emission can leave pending plus EK runs.
template <int W, int EK = 1>
struct OutFifo {
  static constexpr int RUN = W * 8;                                     // bytes per step
  static constexpr int SLACK = kGranule - cgcd(RUN, kGranule);          // most bytes left pending after an emission
  static constexpr int CB = (SLACK + EK * RUN + 15) / 16 * 16;          // FIFO bytes
  static constexpr bool kEven = (W % 2 == 0);                           // every run boundary is 16-byte aligned
  uint32_t base;   // shared-memory address of this system's FIFO (16-byte aligned)
  int off;         // (global byte offset of the current run) mod CB
  int pend;        // bytes at and above the current run that are in the FIFO but not yet stored
  __device__ __forceinline__ void init(uint32_t smem_base, long g_top) {
    base = smem_base;
    off = (int)((g_top * RUN) % CB);   // g_top = global run index one past the trajectory's last run
    pend = 0;
  }
  __device__ __forceinline__ void advance() {   // to the next (lower) run: once per time index, before put
    off -= RUN;
    if (off < 0) off += CB;
    pend += RUN;
  }
  template <int CNT>
  __device__ __forceinline__ void put_at(int e0, const double *v) const {   // doubles e0 .. e0+CNT-1 of the run
    int o = off + 8 * e0;
    if (o >= CB) o -= CB;
    const int room = CB - o;
#pragma unroll
    for (int i = 0; i < CNT; i++) {
      const int d = 8 * i >= room ? 8 * i - CB : 8 * i;
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + (uint32_t)(o + d)), "d"(v[i]) : "memory");
    }
  }
  // bytes [gs, gs + len) of the array (inside one granule) from FIFO offset so
  __device__ __forceinline__ void copy_out(char *out, long gs, int so, int len) const {
    if constexpr (!kEven) {   // run boundaries are only 8-byte aligned: ragged edges by plain stores
      if (gs & 8) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + (uint32_t)so));
        *reinterpret_cast<double *>(out + gs) = v;
        gs += 8;
        so += 8;
        if (so >= CB) so -= CB;
        len -= 8;
      }
      if (len & 8) {
        int st = so + len - 8;
        if (st >= CB) st -= CB;
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + (uint32_t)st));
        *reinterpret_cast<double *>(out + gs + len - 8) = v;
        len -= 8;
      }
      if (len <= 0) return;
    }
    const int l1 = len < CB - so ? len : CB - so;
    bulk_s2g(out + gs, base + (uint32_t)so, (uint32_t)l1);
    if (l1 < len) bulk_s2g(out + gs + l1, base, (uint32_t)(len - l1));
  }
  // After the current run (global run index g, time index t) is in the FIFO: store every granule it
  // completed (everything at t = 0).  Called by ONE lane of the system; the caller commits the group.
  __device__ __forceinline__ void emit(char *out, long g, int t) {
    const long a = g * RUN;
    const int d = (t == 0) ? 0 : (int)((kGranule - (a & (kGranule - 1))) & (kGranule - 1));   // up to the next boundary
    if (pend > d) {
      long gs = a + d;
      int so = off + d;
      if (so >= CB) so -= CB;
      int len = pend - d;
      while (len > 0) {
        const int room = kGranule - (int)(gs & (kGranule - 1));
        const int l = len < room ? len : room;
        copy_out(out, gs, so, l);
        gs += l;
        so += l;
        if (so >= CB) so -= CB;
        len -= l;
      }
      pend = d;
    }
  }
};

template <class Fam, int CK, int U>
struct TpxLayout {
  static constexpr int p = Fam::p, nb = Fam::nb, n = Fam::n, PS = Fam::PS;
  static constexpr int SPW = 32 / nb, LANES = SPW * nb;
  // Draws are 8 n bytes per step (Lorenz63: 72): a granule completes every 256 / (8 n) steps per system.  The
  // emission round (proxy fence, warp barriers, per-lane bulk-copy issue loop) runs every EK steps instead of
  // every step: ncu showed it was ~40 % of the kernel's instructions at EK = 1.
  static constexpr int EK = 4;
  using Fifo = OutFifo<n, EK>;
  // per-system stride: an odd number of 16-byte chunks (FIFO rows of different systems start in different banks)
  static constexpr int SYS_BYTES = (Fifo::CB / 16) % 2 == 1 ? Fifo::CB : Fifo::CB + 16;
  static constexpr int WARPS = 2, THREADS = WARPS * 32;
  // U = 1: 12 resident warps per SM (<= 170 registers) hold a whole 16,384-system Lorenz63 batch (11.1 warps per
  // SM) in ONE wave; at the 190 registers ptxas would take, 10 fit and the launch needs two
  static constexpr int MINB = U == 1 ? 6 : 1;
#ifndef RODEO_TPX_PLAIN
#define RODEO_TPX_PLAIN 1
#endif
  // Output path.  kPlain: every step the lanes park their block's p draws in a per-warp shared-memory image
  // [system][n] (double buffered: one warp barrier per step) and the warp writes it back with plain coalesced
  // stores, consecutive lanes = consecutive doubles of a system's 8n-byte run.  The kernel is latency-bound, not
  // HBM-bound (ncu: DRAM 41 %), and the per-system FIFOs + per-lane bulk copies of aligned granules cost more
  // instructions (serialised UBLKCP issue loops, proxy fences, wait_group) than the partial sectors cost bandwidth.
  static constexpr bool kPlain = RODEO_TPX_PLAIN != 0;
  static constexpr int IMG = SPW * n;                       // doubles of one staging image
  static constexpr int NST = (IMG + 31) / 32;               // store instructions per step
  static constexpr size_t out_bytes = kPlain ? (size_t)WARPS * U * 2 * IMG * 8 : (size_t)WARPS * U * SPW * SYS_BYTES;
#ifndef RODEO_TPX_ASYNC
#define RODEO_TPX_ASYNC 1
#endif
  // History prefetch.  kAsync: each lane copies its block's next checkpoint record (REC doubles) into its own
  // shared-memory slots with cp.async (LDGSTS) one group ahead and picks it up with cp.async.wait_group + LDS.
  // With register prefetches (plain loads a group ahead) the in-flight DRAM loads share a scoreboard with the
  // z loads of the current step: the DFMA that consumes z waited for the PREFETCH to land (ncu: 27 % of all stall
  // samples on that one instruction, unchanged when the z loads were hoisted to the top of the step).
  static constexpr bool kAsync = RODEO_TPX_ASYNC != 0;
  static constexpr int REC = p + PS;
  static constexpr size_t stage_off = (out_bytes + 15) / 16 * 16;
  // Right-hand sides with more than three parameters (MSEIR: six) park theta in per-lane shared-memory slots and
  // re-read it in the re-run step: held in registers it pushes the five-block instance over the 168-register cap
  // (44 bytes of spills, solve_sim smoother 3.32 -> 3.61 ms with the U-form draw).
  static constexpr int NTH = Fam::Ode::n_theta;
  static constexpr bool kThetaSmem = CK > 1 && NTH > 3;
  static constexpr size_t theta_off = stage_off + (kAsync ? (size_t)THREADS * U * 2 * REC * 8 : 0);
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

  long b[U], bs[U], g[U];
  bool valid[U];
  typename L::Fifo fx[U];
  const double *hb[U], *zb[U];
  double th[U][NT], x0b[U][p];
#pragma unroll
  for (int u = 0; u < U; u++) {
    b[u] = b_first + u * SPW + s;
    valid[u] = lane_on && b[u] < a.B;
    bs[u] = b[u] < a.B ? b[u] : b_first;                   // a safe system for the loads of idle lanes
    g[u] = (bs[u] + 1) * T1;
    fx[u].init(smem_u32(smem_raw) + (uint32_t)(((warp * U + u) * SPW + s) * L::SYS_BYTES), g[u]);
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
  // (kAsync) cp.async prefetch of record r into stage r & 1 / pick-up from it; slot of element e: [u][stage][e][thread]
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
  // (kPlain) staging images (shared-state addresses) and this lane's store offsets: double k * 32 + lane of an image =
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
      ov[u][k] = L::kPlain && idx < L::IMG && b_first + u * SPW + sy < a.B;
      ooff[u][k] = (int)((u * SPW + sy) * T1 * n) + el;
    }
  // park the draws of time index t in the FIFOs; every EK steps (and at t = 0) store the granules they completed
  int since = 0;
  bool reads_pending = false;
  auto emit = [&](int t, double (*x)[p]) {
    if constexpr (L::kPlain) {   // double buffered: the reads of the other image precede this step's barrier
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
      return;
    }
    if (reads_pending) {   // the bulk copies of the last emission round have left the FIFOs
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      reads_pending = false;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      g[u] -= 1;
      fx[u].advance();
      fx[u].template put_at<p>(blk * p, x[u]);
    }
    if (++since == L::EK || t == 0) {
      since = 0;
      fence_proxy_async();   // the rows (generic proxy) become visible to the bulk copies (async proxy) ...
      __syncwarp();          // ... of every lane of a system, before the lane of block 0 issues the stores
#pragma unroll
      for (int u = 0; u < U; u++)
        if (valid[u] && blk == 0) fx[u].emit(reinterpret_cast<char *>(a.x_out), g[u], t);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      reads_pending = true;
    }
  };
  // z[:, t-1], this block's rows.  Volatile loads: issued HERE, at the top of the step; left to the compiler they sink
  // to their use at the end of the step (x = m~ + L z) and their L2 latency lands on the chain -- ncu: 27 % of
  // all stall samples on that one DFMA (profiles/r02_backward_tpx_plain_stores_ncu_summary.txt)
  auto z_at = [&](int u, int t, double *zt) {
#pragma unroll
    for (int e = 0; e < p; e++) zt[e] = ld_once(zb[u] + (long)(t - 1) * n + e);
  };

  double x[U][p], cm[U][p], cS[U][PS], nm[U][p], nS[U][PS], mus_unused[p], Ss_unused[PS];
  int fail[U];
#pragma unroll
  for (int u = 0; u < U; u++) fail[u] = 0;
#pragma unroll
  for (int u = 0; u < U; u++) {
    load_rec(u, 0, cm[u], cS[u]);
    if (R > 1) {
      if constexpr (L::kAsync) prefetch_rec(u, 1);
      else load_rec(u, 1, nm[u], nS[u]);
    }
  }
  if constexpr (L::kAsync) asm volatile("cp.async.commit_group;" ::: "memory");
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
    if (base >= 1) {
      if constexpr (L::kAsync) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");          // record r (copied a group ago) is in its stage
#pragma unroll
        for (int u = 0; u < U; u++) {
          fetch_rec(u, r, cm[u], cS[u]);
          if (r + 1 < R) prefetch_rec(u, r + 1);                      // one group ahead, into the other stage
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      } else {
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
          for (int e = 0; e < p; e++) cm[u][e] = nm[u][e];
#pragma unroll
          for (int e = 0; e < PS; e++) cS[u][e] = nS[u][e];
          if (r + 1 < R) load_rec(u, r + 1, nm[u], nS[u]);            // one group ahead
        }
      }
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
          double fm[p], fS[PS], zt[p];
          z_at(u, t, zt);
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
  if constexpr (!L::kPlain) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
