// kalman_tpb.cuh -- backward (smoother) kernel of the block-diagonal family with one THREAD PER
// DIAGONAL BLOCK and granule-aligned output through TMA tensor stores (round 2; replaces
// kalman_staged.cuh for solve_mv / solve_sim / solve on two-block priors: the FitzHugh-Nagumo headline).
//
// Why (measured on B200, tools/microbench/write_pattern3.cu, FN solve_mv output = 8.83 GB):
//   one 288-byte variance run + one 48-byte mean run per system and step (round 1)      2.43 ms  3.6 TB/s
//   the same runs as TMA tensor / bulk stores, unaligned                                  2.5 ms
//   ONLY whole 256-byte-aligned granules, each written once by the TMA                   1.54 ms  5.75 TB/s
//   ... with the checkpoint-2 history read stream beside it                              1.93 ms  5.5 TB/s
// The L2 slice hash uses address bits >= 8, so a 256-byte aligned granule lives in one slice and reaches
// DRAM as one burst; a run that straddles granules reaches it as two partial writes at different times.
//
// Mapping: lane = blk * 16 + i; warp (q, j) owns the 16 systems b_i = 256 q + 16 i + j.  Sixteen systems
// apart the trajectories are congruent modulo 256 bytes (16 * 8 * n doubles per step is a multiple of 256 for
// even n), so the whole warp completes its output granules at the same steps: the bookkeeping is
// warp-uniform and ONE cp.async.bulk.tensor serves all 16 systems.  The forward pass writes the history in
// the matching permuted order (SolveArgs::perm16), so the warp's 16 systems are contiguous there.
//   * the blocks of a system decouple in the smoother (kalman_common.cuh); only the ODE right-hand side
//     couples them, in the re-run filter step of the checkpointed history: one shuffle per block.  Against one
//     thread per system this doubles the resident warps per shared-memory byte and halves the FP64 chain.
//   * history: each lane prefetches its block's p + p(p+1)/2 doubles of the NEXT checkpoint record into
//     registers one group (CK steps) ahead (half-warp = 128 contiguous bytes per element row); no ring.
//   * checkpoint interval CK in {1, 2}: for CK = 2 the filter is re-run from the checkpoint for one step in
//     registers (same routines as the forward pass: bit-identical moments).
//   * outputs: per output array the warp keeps NSLOT granule slots in shared memory, a slot being ONE TMA box
//     image {16 doubles, 2, 16 systems} (4 KB: [system][half][128 bytes], 128-byte swizzle: chunk c of 128-byte
//     row r = 2 i + half sits at chunk c ^ (r & 7), a 2-way bank conflict on the per-lane row writes, which is
//     cheap; one store per whole granule is what the memory system wants).  A slot holds granule (x >> 8) % NSLOT
//     of every system, x = byte offset inside the 16-system row.  Every lane writes its block's rows of the
//     step's run (structural zeros included); when the run completes a granule, lane 0 issues one tensor
//     store.  Trajectory ends (partial granules) and the systems of a last incomplete 16-system row are
//     stored with plain 16-byte stores.
// Warps never synchronise with each other.
#pragma once
#include <cstdint>
#include <cstdlib>

#include "kalman_staged.cuh"

namespace rodeo {

constexpr int kGranule = 256;
constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }

// Tensor map of one output array for the granule stores: W doubles per (system, time index).
// Byte address = k * PITCH + x with k = system / 16 and x the offset inside the 16-system row,
// PITCH = 16 * T1 * W * 8.  Dimensions {16 doubles, x / 128, k < B / 16}, box {16, G / 128, 16}: one whole granule
// of 16 systems per store, [system][G / 128][128 bytes] in shared memory.  ONE store per granule matters: with two
// 128-byte-wide boxes per 256-byte granule the halves of a granule reach L2 / DRAM separated in time and the
// memory system runs at the 128-byte-granule rate (write_pattern3: 4.6 vs 5.75 TB/s) -- measured in the kernel:
// 1.97 ms with box {16, 1, 16}, see NOTES_r02.md for the one-store figure.
inline cudaError_t encode_granule_map(CUtensorMap *m, double *base, int W, long T1, long B, int G = kGranule) {
  memset(m, 0, sizeof *m);
  if (!base || B < 16) return cudaSuccess;     // no complete 16-system row: only plain stores are used
  encode_tiled_fn enc = tensor_map_encoder();
  if (!enc) return cudaErrorNotSupported;
  const cuuint64_t pitch = (cuuint64_t)16 * T1 * W * 8;
  const cuuint64_t dims[3] = {16, pitch / 128, (cuuint64_t)(B / 16)};
  const cuuint64_t strides[2] = {128, pitch};
  const cuuint32_t box[3] = {16, (cuuint32_t)(G / 128), 16};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// shared -> global tensor store of one box, source given as a shared-state-space address
__device__ __forceinline__ void tma_store_3d_s(const CUtensorMap *map, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

struct GranuleMaps {
  alignas(64) CUtensorMap mu, var, x;
};

// One output stream (W doubles per system and step) of one warp.  All members are warp-uniform except
// tile_i / i7s / out_i (the lane's system).
// G = granule bytes (a power of two >= 256 with 16 * RUN * T1 a multiple of G for every T1: 256 always works for
// even W; 512 when 16 * RUN is a multiple of 512, e.g. the 288-byte variance runs of n_state 6).
template <int W, int G = kGranule>
struct GranuleStream {
  static constexpr int RUN = W * 8;                                  // bytes per system and step
  static constexpr int kG = G;
  static_assert(G >= 256 && (G & (G - 1)) == 0 && (16 * W * 8) % G == 0, "granule size");
  static constexpr int SLACK = G - cgcd(RUN, G);       // most bytes left pending after an emission
  static constexpr int NSLOT = (SLACK + RUN + G - 1) / G;
  static constexpr int UNITS = NSLOT * (G / 128);                    // 128-byte units per system in the slots
  static constexpr int BYTES = NSLOT * 16 * G;
  static_assert(W % 2 == 0, "granule stores need 16-byte run boundaries (even n_state)");
  // row coordinates are 32-bit: the launcher checks 16 * T1 * RUN < 2^31
  int x;           // offset of the current run inside the 16-system row (bytes): (j * T1 + t) * RUN
  int Fx;          // everything at or above Fx (a granule boundary) is stored or not ours
  int x_end;       // end of this warp's trajectories inside the row: (j + 1) * T1 * RUN
  uint32_t tile;   // shared address of the slots (1024-byte aligned)
  uint32_t tile_i; // tile + i * G: this lane's system inside a slot
  uint32_t r0;     // i * (G / 128): index of this lane's first 128-byte row inside a slot (the swizzle key)
  char *out_i;     // array + (system / 16) * PITCH: this lane's row base in global memory
  __device__ __forceinline__ void init(uint32_t tile_, int i, double *out, long k_row, int j, long T1) {
    x_end = (int)((j + 1) * T1 * RUN);
    x = x_end;
    Fx = (x_end + G - 1) & ~(G - 1);
    tile = tile_;
    tile_i = tile + (uint32_t)i * (uint32_t)G;
    r0 = (uint32_t)i * (uint32_t)(G / 128);
    out_i = reinterpret_cast<char *>(out) + k_row * (16 * T1 * RUN);
  }
  __device__ __forceinline__ void advance() { x -= RUN; }
  // shared address of byte xx (a multiple of 8) of this lane's system
  __device__ __forceinline__ uint32_t phys(int xx) const {
    const uint32_t u = (uint32_t)xx;
    // slot image = the TMA box {16 doubles, G / 128, 16 systems} with the 128-byte swizzle: 128-byte row
    // r = i * (G / 128) + hh of the slot, its 16-byte chunk c stored at chunk c ^ (r & 7)
    const uint32_t unit = (u >> 7) % UNITS, hh = unit % (G / 128), slot = unit / (G / 128);
    const uint32_t swz = ((r0 + hh) & 7u) << 4;
    return tile_i + slot * (16u * G) + hh * 128u + (((u & 0x70u) ^ swz) | (u & 8u));
  }
  // CNT doubles of the current run starting at byte offset rb (lane dependent) of the run
  template <int CNT>
  __device__ __forceinline__ void put_pairs(int rb, const double *v) const {
    const int xb = x + rb;
#pragma unroll
    for (int k = 0; k < CNT; k += 2)
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(phys(xb + 8 * k)), "d"(v[k]), "d"(v[k + 1]) : "memory");
  }
  template <int CNT>
  __device__ __forceinline__ void put_singles(int rb, const double *v) const {
    const int xb = x + rb;
#pragma unroll
    for (int k = 0; k < CNT; k++) asm volatile("st.shared.f64 [%0], %1;" ::"r"(phys(xb + 8 * k)), "d"(v[k]) : "memory");
  }
  // bytes [xa, xb) of this lane's system with plain 16-byte loads / stores (generic proxy)
  __device__ __forceinline__ void copy_lsu(int xa, int xb) const {
    for (int xx = xa; xx < xb; xx += 16) {
      double v0, v1;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(phys(xx)));
      *reinterpret_cast<double2 *>(out_i + xx) = make_double2(v0, v1);
    }
  }
  // After the run of time index t is in the slots (and fenced): store what it completed.
  //   full_row: this warp has systems in complete 16-system rows (lane 0 issues the tensor stores)
  //   mine:     this lane stores its system's partial granules (one lane per valid system)
  //   tail_row: this lane's system sits in the last, incomplete row: it stores whole granules itself too
  __device__ __forceinline__ void emit(const CUtensorMap *map, int t, int row0, bool full_row, bool lane0, bool mine,
                                       bool tail_row) {
    const int c = (t == 0) ? x : (x + G - 1) & ~(G - 1);
    for (int g0 = (c + G - 1) & ~(G - 1); g0 < Fx; g0 += G) {
      if (g0 + G > x_end) {                     // the top granule is shared with the next trajectory
        if (mine) copy_lsu(g0, x_end);
      } else {
        if (full_row && lane0) {
          tma_store_3d_s(map, tile + (((uint32_t)g0 / G) % NSLOT) * (16u * G), 0, g0 >> 7, row0);
        }
        if (mine && tail_row) copy_lsu(g0, g0 + G);
      }
    }
    if (t == 0 && (c & (G - 1))) {              // the bottom granule is shared with the previous trajectory
      const int top = (c + G - 1) & ~(G - 1);
      if (mine) copy_lsu(c, top < x_end ? top : x_end);
    }
    Fx = (c + G - 1) & ~(G - 1);
  }
};

// DO_VAR: 0 no variance output, 1 the reference's full (n, n) matrices, 2 the n_meas diagonal (p, p) blocks only
// (RODEO_VAR_BLOCKS), 3 the marginal variances only (RODEO_VAR_DIAG) -- rodeo_cuda_set_var_layout
template <class Fam, int CK, bool DO_MEAN, int DO_VAR, bool DO_SIM>
struct TpbLayout {
  static constexpr int p = Fam::p, nb = Fam::nb, n = Fam::n, PS = Fam::PS;
  static_assert(nb == 2, "thread-per-block smoother: two diagonal blocks (16 systems x 2 blocks per warp)");
  static_assert(CK == 1 || CK == 2 || CK == 4, "thread-per-block smoother: checkpoint interval 1, 2 or 4");
#ifndef RODEO_TPB_GVAR
#define RODEO_TPB_GVAR 256
#endif
  // variance granules of RODEO_TPB_GVAR bytes where the run length allows it (see GranuleStream), else 256
  static constexpr int GVAR = (16 * n * n * 8) % RODEO_TPB_GVAR == 0 ? RODEO_TPB_GVAR : kGranule;
  static constexpr int WV = DO_VAR == 2 ? nb * p * p : DO_VAR == 3 ? n : n * n;   // variance doubles per system and step
  using SVec = GranuleStream<n>;
  using SVar = GranuleStream<WV, DO_VAR == 1 ? GVAR : kGranule>;
  static constexpr int OFF_MEAN = 0;
  static constexpr int OFF_X = OFF_MEAN + (DO_MEAN ? SVec::BYTES : 0);
  static constexpr int OFF_VAR = OFF_X + (DO_SIM ? SVec::BYTES : 0);
  static constexpr int WARP_BYTES = OFF_VAR + (DO_VAR ? SVar::BYTES : 0);      // a multiple of 4096
#ifndef RODEO_TPB_WARPS
#define RODEO_TPB_WARPS 2
#endif
  static constexpr int WARPS = RODEO_TPB_WARPS;
  static constexpr int THREADS = WARPS * 32;
#ifndef RODEO_TPB_ASYNC
#define RODEO_TPB_ASYNC 1
#endif
  // Variants that draw (DO_SIM) read z[:, t-1] inside every step.  With the history prefetched into REGISTERS a
  // group ahead, those z loads share a scoreboard with the in-flight DRAM loads and their consumer waits for the
  // prefetch (found on the Lorenz63 draw smoother, kalman_tpx.cuh).  They prefetch with cp.async into per-lane
  // shared-memory slots instead and pin the z loads at the top of the step.
  // Draw-only variants: with the variance slots beside it the extra shared memory costs a resident CTA
  // (FN solve: 2.80 -> 3.09 ms), and the log-likelihood kernels have no dynamic shared memory at all.
  static constexpr bool kAsync = DO_SIM && DO_VAR == 0 && RODEO_TPB_ASYNC != 0;
  static constexpr int REC = p + PS;
  static constexpr int STAGE_OFF = WARPS * WARP_BYTES;                         // after the output slots
  static constexpr size_t smem_need = (size_t)WARPS * WARP_BYTES + (kAsync ? (size_t)THREADS * 2 * REC * 8 : 0) +
                                      1024;                                    // + alignment slack (1024-byte swizzle atoms)
  // The kernel is HBM-bound and runs best with FEWER resident warps (measured, FN solve_mv, 65,536 systems:
  // 12 warps per SM 2.05 ms, 10: 2.00, 8: 1.99, 6: 1.98, 4: 2.40): with 4,096 equal-length warps the last,
  // partially filled wave is a smaller share of the launch.  Requesting 55 KB per CTA caps it at 4 CTAs = 8 warps.
  // Only the variants that write the full variance are HBM-bound; the draw-only and compact-layout variants are
  // latency / FP64-bound and take every warp the registers allow.
  static constexpr size_t smem_bytes = (DO_VAR == 1 && smem_need < 55 * 1024) ? 55 * 1024 : smem_need;
};

// The prior of the lane's block: immediate constant-bank operands when every block has the same Q, R, w, c
// (what ibm_init with equal sigma / n_deriv_prior gives, every BASELINE config), per-lane register copies otherwise.
template <class Fam, bool UNIFORM>
struct BlockPrior;
template <class Fam>
struct BlockPrior<Fam, true> {
  const typename Fam::PriorT &P;
  __device__ __forceinline__ BlockPrior(const typename Fam::PriorT &P_, int) : P(P_) {}
  __device__ __forceinline__ const double *Q() const { return P.Q[0]; }
  __device__ __forceinline__ const double *R() const { return P.R[0]; }
  __device__ __forceinline__ const double *w() const { return P.w[0]; }
  __device__ __forceinline__ const double *c() const { return P.c; }
};
template <class Fam>
struct BlockPrior<Fam, false> {
  static constexpr int p = Fam::p, PS = Fam::PS;
  double q_[p * p], r_[PS], w_[p], c_[p];
  __device__ __forceinline__ BlockPrior(const typename Fam::PriorT &P, int blk) {
#pragma unroll
    for (int i = 0; i < p * p; i++) q_[i] = P.Q[blk][i];
#pragma unroll
    for (int i = 0; i < PS; i++) r_[i] = P.R[blk][i];
#pragma unroll
    for (int i = 0; i < p; i++) {
      w_[i] = P.w[blk][i];
      c_[i] = P.c[blk * p + i];
    }
  }
  // per-system priors (rodeo_cuda_solve_batched_prior): overlay system b's own block of every array that is
  // given (device arrays, batch-major, each system column-major like the constructor's; nullptr = shared value)
  __device__ __forceinline__ void overlay(const BatchedPrior &bq, long b, int blk) {
    constexpr int n = Fam::n, m = Fam::m;
    if (bq.Q) {
      const double *q = bq.Q + b * n * n;
#pragma unroll
      for (int i = 0; i < p; i++)
#pragma unroll
        for (int j = 0; j < p; j++) q_[i * p + j] = q[(blk * p + i) + (blk * p + j) * n];
    }
    if (bq.R) {
      const double *r = bq.R + b * n * n;
#pragma unroll
      for (int i = 0; i < p; i++)
#pragma unroll
        for (int j = i; j < p; j++) r_[sidx<p>(i, j)] = r[(blk * p + i) + (blk * p + j) * n];
    }
    if (bq.W) {
      const double *w = bq.W + b * m * n;
#pragma unroll
      for (int i = 0; i < p; i++) w_[i] = w[blk + (blk * p + i) * m];
    }
    if (bq.c) {
#pragma unroll
      for (int i = 0; i < p; i++) c_[i] = bq.c[b * n + blk * p + i];
    }
  }
  __device__ __forceinline__ const double *Q() const { return q_; }
  __device__ __forceinline__ const double *R() const { return r_; }
  __device__ __forceinline__ const double *w() const { return w_; }
  __device__ __forceinline__ const double *c() const { return c_; }
};

// PM, the prior mode of the thread-per-block kernels: 0 = every block has the same Q, R, w, c (immediate
// constant-bank operands), 1 = per-lane register copies of the lane's block, 2 = per-SYSTEM priors overlaid on them
__device__ __forceinline__ double ld_once(const double *p) {
  double v;   // volatile: loaded where it is written and kept (never re-loaded inside the time loop)
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;   // volatile: keeps the prefetch where it is written (a group ahead of its use)
  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// LL: no per-step output; the thinned Gaussian log-likelihood of the smoothed mean (DO_MEAN) or of the draw (DO_SIM)
// is accumulated per system (examples/inference.py:54-56, 66-94) -- no shared memory, no stores but one double.
template <class Fam, int CK, int PM, bool DO_MEAN, int DO_VAR, bool DO_SIM, bool LL = false>
__global__ void __launch_bounds__(TpbLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>::THREADS)
#ifdef RODEO_TPB_MAXNREG
__maxnreg__(RODEO_TPB_MAXNREG)
#endif
kalman_backward_tpb(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a,
                    const __grid_constant__ GranuleMaps maps, const __grid_constant__ BatchedPrior bpr) {
  using L = TpbLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>;
  using Ode = typename Fam::Ode;
  constexpr int p = L::p, nb = L::nb, n = L::n, PS = L::PS, NE = Fam::NE;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr bool kAsyncK = L::kAsync && !LL;   // cp.async history prefetch (TpbLayout)
  extern __shared__ unsigned char smem_raw[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long wg = (long)blockIdx.x * L::WARPS + warp;     // warp index in the grid
  const long q = wg >> 4;
  const int j = (int)(wg & 15);
  if (256 * q + j >= a.B) return;                          // warp-uniform; warps are independent
  const int blk = lane >> 4, i = lane & 15;
  const long b = 256 * q + 16 * i + j;                     // this lane's system
  const bool valid = b < a.B;
  const long bs = valid ? b : 256 * q + j;                 // a safe system for the loads of idle lanes
  const long upos = 256 * q + 16 * j + i;                  // its position in the (permuted) history
  const int N = P.n_eval, R = n_records(N, CK);
  const long T1 = (long)N + 1;
  BlockPrior<Fam, PM == 0> bp(P, blk);
  if constexpr (PM == 2) bp.overlay(bpr, bs, blk);

  // ---- output streams -------------------------------------------------------------------------------
  const uint32_t tile = ((smem_u32(smem_raw) + 1023u) & ~1023u) + (uint32_t)warp * L::WARP_BYTES;
  const long k_row = 16 * q + i;                           // b / 16
  const long full_rows = a.B / 16;
  const bool full_row = 16 * q < full_rows;                // the tensor stores of this warp have rows to write
  const bool tail_row = k_row >= full_rows;                // (valid lanes) stored with plain stores only
  const bool mine = valid && blk == 0, lane0 = lane == 0;
  typename L::SVec sm, sx;
  typename L::SVar sv;
  if (!LL && DO_MEAN) sm.init(tile + L::OFF_MEAN, i, a.mu_out, k_row, j, T1);
  if (!LL && DO_SIM) sx.init(tile + L::OFF_X, i, a.x_out, k_row, j, T1);
  if (!LL && DO_VAR) sv.init(tile + L::OFF_VAR, i, a.var_out, k_row, j, T1);
  double lp = 0.0;                 // (LL) this system's log-likelihood so far
  int d_obs = a.n_obs_times - 1;   // (LL) next observation time, walking backwards

  auto emit = [&](int t, const double *mus, const double *Ss, const double *x) {
    if constexpr (LL) {
      // At an observed time index the lanes of a system pool the state vector and add the terms in the order of
      // the thread-per-system kernels (LoglikEmitter, kalman_staged.cuh), so the sums are bit-identical to theirs.
      while (d_obs >= 0 && a.thin_index[d_obs] == t) {       // warp-uniform
        double X[n];
#pragma unroll
        for (int k = 0; k < nb; k++)
#pragma unroll
          for (int e = 0; e < p; e++) X[k * p + e] = __shfl_sync(kFull, DO_SIM ? x[e] : mus[e], k * 16 + i);
        lp += reg::loglik_terms<n>(a, d_obs, X);
        d_obs--;
      }
      return;
    }
    if (lane0) tma_wait_read();   // the earlier tensor stores have left the slots
    __syncwarp();
    if (DO_MEAN) {
      sm.advance();
      sm.template put_singles<p>(blk * p * 8, mus);
    }
    if (DO_SIM) {
      sx.advance();
      sx.template put_singles<p>(blk * p * 8, x);
    }
    if (DO_VAR) {
      sv.advance();
      if constexpr (DO_VAR == 1) {
        // rows blk*p .. blk*p+p-1 of the n x n matrix (the reference's var[t][j][i]; symmetric), zeros outside the block
#pragma unroll
        for (int jj = 0; jj < p; jj++) {
          double row[n];
#pragma unroll
          for (int c = 0; c < n; c++) row[c] = (blk == c / p) ? Ss[sidx<p>(c % p, jj)] : 0.0;
          sv.template put_pairs<n>((blk * p + jj) * n * 8, row);
        }
      } else if constexpr (DO_VAR == 2) {   // this block's p x p matrix, var[t][blk][jj][ii]
        double bv[p * p];
#pragma unroll
        for (int jj = 0; jj < p; jj++)
#pragma unroll
          for (int ii = 0; ii < p; ii++) bv[jj * p + ii] = Ss[sidx<p>(ii, jj)];
        if constexpr ((p * p) % 2 == 0) sv.template put_pairs<p * p>(blk * p * p * 8, bv);
        else sv.template put_singles<p * p>(blk * p * p * 8, bv);
      } else {                              // marginal variances of this block's components
        double dv[p];
#pragma unroll
        for (int ii = 0; ii < p; ii++) dv[ii] = Ss[sidx<p>(ii, ii)];
        sv.template put_singles<p>(blk * p * 8, dv);
      }
    }
    fence_proxy_async();   // the rows (generic proxy) become visible to the tensor stores (async proxy) ...
    __syncwarp();          // ... of every lane, before lane 0 issues them
    const int row0 = (int)(16 * q);
    if (DO_VAR) sv.emit(&maps.var, t, row0, full_row, lane0, mine, tail_row);
    if (DO_MEAN) sm.emit(&maps.mu, t, row0, full_row, lane0, mine, tail_row);
    if (DO_SIM) sx.emit(&maps.x, t, row0, full_row, lane0, mine, tail_row);
    if (lane0) tma_commit();
  };

  const double *zb_ = DO_SIM ? a.z + bs * a.z_stride + blk * p : nullptr;
  // ---- history: this block's elements of record r --------------------------------------------------------
  const double *hb = a.hist + hist_offset(upos, 0, R, NE);
  auto load_rec = [&](int r, double *mu, double *S) {
    const double *h = hb + (size_t)r * NE * kTile;
#pragma unroll
    for (int e = 0; e < p; e++) mu[e] = ld_stream(h + (blk * p + e) * kTile);
#pragma unroll
    for (int e = 0; e < PS; e++) S[e] = ld_stream(h + (n + blk * PS + e) * kTile);
  };

  // (kAsync) cp.async prefetch of record r into stage r & 1 / pick-up from it; slot of element e: [stage][e][thread]
  const uint32_t stage_base = ((smem_u32(smem_raw) + 1023u) & ~1023u) + (uint32_t)L::STAGE_OFF + (uint32_t)threadIdx.x * 8u;
  auto stage_addr = [&](int st, int e) { return stage_base + (uint32_t)(((st * L::REC + e) * L::THREADS) * 8); };
  auto prefetch_rec = [&](int r) {
    const double *h = hb + (size_t)r * NE * kTile;
    const int st = r & 1;
#pragma unroll
    for (int e = 0; e < p; e++)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stage_addr(st, e)), "l"(h + (blk * p + e) * kTile) : "memory");
#pragma unroll
    for (int e = 0; e < PS; e++)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stage_addr(st, p + e)), "l"(h + (n + blk * PS + e) * kTile)
                   : "memory");
  };
  auto fetch_rec = [&](int r, double *mu, double *S) {
    const int st = r & 1;
#pragma unroll
    for (int e = 0; e < p; e++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(mu[e]) : "r"(stage_addr(st, e)) : "memory");
#pragma unroll
    for (int e = 0; e < PS; e++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(S[e]) : "r"(stage_addr(st, p + e)) : "memory");
  };
  // z[:, t-1], this block's rows, loaded where this is called (volatile), not where the draw consumes it
  auto z_at = [&](int t, double *zt) {
#pragma unroll
    for (int e = 0; e < p; e++) zt[e] = DO_SIM ? ld_once(zb_ + (long)(t - 1) * n + e) : 0.0;
  };

  // theta and x0 are loaded ONCE with volatile loads: left to the compiler, the x0 load of the last group
  // (base < 1) is speculated into every group and its L2 latency lands on the critical path
  double th[NT], x0b[p];
#pragma unroll
  for (int e = 0; e < p; e++) x0b[e] = ld_once(a.x0 + bs * n + blk * p + e);
#pragma unroll
  for (int e = 0; e < NT; e++) th[e] = (Ode::n_theta > 0 && CK > 1) ? ld_once(a.theta + bs * Ode::n_theta + e) : 0.0;

  double mus[p], Ss[PS], x[p];
  double cm[p], cS[PS], nm[p], nS[PS];     // current / prefetched checkpoint record (this block)
  int fail = 0;
  load_rec(0, cm, cS);
  if (R > 1) {
    if constexpr (kAsyncK) prefetch_rec(1);
    else load_rec(1, nm, nS);
  }
  if constexpr (kAsyncK) asm volatile("cp.async.commit_group;" ::: "memory");
  {  // terminal time index N = record 0
    double zt[p];
    z_at(N, zt);
    fail = reg::smooth_terminal<p, 1, DO_MEAN, (DO_VAR != 0), DO_SIM>(cm, cS, zt, mus, Ss, x);
    emit(N, mus, Ss, x);
  }
  int r = 1;
  for (int t_hi = N; t_hi > 1;) {   // this group smooths time indices t_hi-1 .. max(t_hi-CK, 1)
    const int base = t_hi - CK;
    // (draw-only variants) z of the group's first time index, t_hi - 1, is loaded BEFORE the group's asynchronous
    // copies: loaded after them its consumer waits for them (kalman_tpx.cuh)
    double ztA[p];
    if constexpr (kAsyncK && CK > 1) z_at(t_hi - 1, ztA);
    if (base >= 1) {
      if constexpr (kAsyncK) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // record r (copied a group ago) is in its stage
        fetch_rec(r, cm, cS);
        if (r + 1 < R) prefetch_rec(r + 1);                    // one group ahead, into the other stage
        asm volatile("cp.async.commit_group;" ::: "memory");
      } else {
#pragma unroll
        for (int e = 0; e < p; e++) cm[e] = nm[e];
#pragma unroll
        for (int e = 0; e < PS; e++) cS[e] = nS[e];
        if (r + 1 < R) load_rec(r + 1, nm, nS);            // one group ahead
      }
    }
    if constexpr (CK > 1) {
      const int t0 = base >= 1 ? base : 0;
      if (base < 1) {   // the known initial state (KalmanODE.pyx:293-297)
#pragma unroll
        for (int e = 0; e < p; e++) cm[e] = x0b[e];
#pragma unroll
        for (int e = 0; e < PS; e++) cS[e] = 0.0;
      }
      const int nfwd = t_hi - 1 - t0;   // time indices t0+1 .. t_hi-1 are re-run from the checkpoint (0 .. CK-1 of them)
      // re-run the filter (BdFam::step, this block's share; the right-hand side needs every block's predicted
      // first component: one shuffle per block), keeping the CK-1 intermediate filtered states in registers
      double sm_[CK - 1][p], sS_[CK - 1][PS];
#pragma unroll
      for (int jf = 0; jf < CK - 1; jf++) {
        if (jf < nfwd) {
          const double *pm = jf == 0 ? cm : sm_[jf == 0 ? 0 : jf - 1];
          const double *pS = jf == 0 ? cS : sS_[jf == 0 ? 0 : jf - 1];
          reg::predict_mean<p, Fam::kUT>(bp.Q(), bp.c(), pm, sm_[jf]);
          reg::predict_var<p, false, Fam::kUT>(bp.Q(), bp.R(), pS, sS_[jf], nullptr);
          double X[n], y[Fam::m];
#pragma unroll
          for (int e = 0; e < n; e++) X[e] = 0.0;
#pragma unroll
          for (int k = 0; k < nb; k++) X[k * p] = __shfl_sync(kFull, sm_[jf][0], k * 16 + i);
          const double tt = P.tmin + (P.tmax - P.tmin) * (t0 + jf + 1) / P.n_eval;
          Ode::template eval<n>(X, tt, th, y);
          double ya = y[0];
#pragma unroll
          for (int k = 1; k < nb; k++) ya = (blk == k) ? y[k] : ya;
          reg::update_block<p, Fam::kWK>(bp.w(), sm_[jf], sS_[jf], ya, false);
        }
      }
#pragma unroll
      for (int jf = CK - 2; jf >= 0; jf--) {
        if (jf < nfwd) {
          const int t = t0 + jf + 1;
          double zt[p];
          if (kAsyncK && t == t_hi - 1) {
#pragma unroll
            for (int e = 0; e < p; e++) zt[e] = ztA[e];
          } else {
            z_at(t, zt);
          }
          const int f = reg::smooth_step<p, 1, DO_MEAN, (DO_VAR != 0), DO_SIM, false, Fam::kUT>(
              bp.Q(), bp.c(), bp.R(), sm_[jf], sS_[jf], zt, mus, Ss, x);
          if (f && !fail) fail = f;
          emit(t, mus, Ss, x);
        }
      }
    }
    if (base >= 1) {
      double zt[p];
      z_at(base, zt);
      const int f = reg::smooth_step<p, 1, DO_MEAN, (DO_VAR != 0), DO_SIM, false, Fam::kUT>(
          bp.Q(), bp.c(), bp.R(), cm, cS, zt, mus, Ss, x);
      if (f && !fail) fail = f;
      emit(base, mus, Ss, x);
    }
    t_hi = base >= 1 ? base : 1;
    r++;
  }
  {  // time index 0: the initial state is known, Sigma_0 = 0 (KalmanODE.pyx:445, 409)
#pragma unroll
    for (int e = 0; e < p; e++) {
      const double v = x0b[e];
      mus[e] = v;
      x[e] = v;
    }
#pragma unroll
    for (int e = 0; e < PS; e++) Ss[e] = 0.0;
    emit(0, mus, Ss, x);
  }
  if constexpr (LL) {
    // constant of the normal density: -(log sqrt(2 pi) + log gamma) per observation (as LoglikEmitter::finish)
    const double cst = -0.91893853320467274178 - log(a.gamma);
    if (valid && blk == 0) a.loglik[b] = lp + cst * (double)(a.n_obs_times * a.n_obs_comp);
  } else {
    if (lane0) tma_wait_all();
  }
  if (valid && a.status && fail) atomicCAS(a.status + b, 0, blk * p + fail);
}

// every block has the same Q, R, w and c?
template <class Fam>
inline bool blocks_uniform(const typename Fam::PriorT &P) {
  constexpr int p = Fam::p, nb = Fam::nb, PS = Fam::PS;
  for (int k = 1; k < nb; k++) {
    for (int i = 0; i < p * p; i++)
      if (P.Q[k][i] != P.Q[0][i]) return false;
    for (int i = 0; i < PS; i++)
      if (P.R[k][i] != P.R[0][i]) return false;
    for (int i = 0; i < p; i++)
      if (P.w[k][i] != P.w[0][i] || P.c[k * p + i] != P.c[i]) return false;
  }
  return true;
}

// Forward filter with the same mapping (lane = blk * 16 + i, warp (q, j) owns systems 256 q + 16 i + j): each lane
// runs its block's predict / update (BdFam::step, same routines: bit-identical moments), the blocks exchange the
// first component of the predicted mean (or of the chkrebtii draw) for the right-hand side with one shuffle
// each, and the filtered records go to the permuted history position 256 q + 16 j + i: a half-warp stores
// 128 contiguous bytes per element row.  Half the dependent FP64 chain per step and twice the warps of the
// thread-per-system kernel (which stays for the unpermuted order).
template <class Fam, int IG, int PM>
__global__ void __launch_bounds__(128)
kalman_forward_tpb(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a,
                   const __grid_constant__ BatchedPrior bpr) {
  using Ode = typename Fam::Ode;
  constexpr int p = Fam::p, nb = Fam::nb, n = Fam::n, PS = Fam::PS, NE = Fam::NE;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  constexpr unsigned kFull = 0xffffffffu;
  static_assert(nb == 2, "thread-per-block forward: two diagonal blocks");
  const int lane = threadIdx.x % 32;
  const long wg = ((long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const long q = wg >> 4;
  const int j = (int)(wg & 15);
  if (256 * q + j >= a.B) return;
  const int blk = lane >> 4, i = lane & 15;
  const long b = 256 * q + 16 * i + j;
  const bool valid = b < a.B;
  const long bs = valid ? b : 256 * q + j;
  const long upos = 256 * q + 16 * j + i;
  BlockPrior<Fam, PM == 0> bp(P, blk);
  if constexpr (PM == 2) bp.overlay(bpr, bs, blk);
  const int N = P.n_eval, ck = a.ck, R = n_records(N, ck);
  double mu[p], S[PS], th[NT];
#pragma unroll
  for (int e = 0; e < p; e++) mu[e] = a.x0[bs * n + blk * p + e];
#pragma unroll
  for (int e = 0; e < PS; e++) S[e] = 0.0;
#pragma unroll
  for (int e = 0; e < NT; e++) th[e] = (Ode::n_theta > 0) ? a.theta[bs * Ode::n_theta + e] : 0.0;
  const double *zb = (IG == IG_CHKREBTII) ? a.z + bs * a.z_stride + blk * p : nullptr;
  double *h = a.hist + hist_offset(upos, R - 1, R, NE);   // records are stored from r = R - 1 down to r = 0
  int t_fail = 0x7fffffff;     // first time index at which this block's innovation variance was not positive
  int back = N - 1;            // distance of time index t + 1 from the end
  int until = back % ck;       // steps until the next stored record
  for (int t = 0; t < N; t++, back--) {
    double mup[p], Sp[PS];
    reg::predict_mean<p, Fam::kUT>(bp.Q(), bp.c(), mu, mup);
    reg::predict_var<p, false, Fam::kUT>(bp.Q(), bp.R(), S, Sp, nullptr);
    double X[n], y[Fam::m], lead = mup[0];
    int f = 0;
    if (IG == IG_CHKREBTII) {   // x_state ~ N(mu_pred, Sigma_pred) with z[:, t]  (KalmanODE.pyx:13-49)
      double xs[p], V[PS], zt[p];
#pragma unroll
      for (int e = 0; e < PS; e++) V[e] = Sp[e];
#pragma unroll
      for (int e = 0; e < p; e++) zt[e] = zb[(long)t * n + e];
      f = reg::state_sim<p>(xs, mup, V, zt);
      lead = xs[0];
    }
#pragma unroll
    for (int e = 0; e < n; e++) X[e] = 0.0;
#pragma unroll
    for (int k = 0; k < nb; k++) X[k * p] = __shfl_sync(kFull, lead, k * 16 + i);
    const double tt = P.tmin + (P.tmax - P.tmin) * (t + 1) / P.n_eval;
    Ode::template eval<n>(X, tt, th, y);
    double ya = y[0];
#pragma unroll
    for (int k = 1; k < nb; k++) ya = (blk == k) ? y[k] : ya;
    const int f2 = reg::update_block<p, Fam::kWK>(bp.w(), mup, Sp, ya, IG == IG_SCHOBERT);
    if ((f | f2) && t_fail == 0x7fffffff) t_fail = t;
#pragma unroll
    for (int e = 0; e < p; e++) mu[e] = mup[e];
#pragma unroll
    for (int e = 0; e < PS; e++) S[e] = Sp[e];
    if (until == 0) {
#pragma unroll
      for (int e = 0; e < p; e++) __stcs(h + (blk * p + e) * kTile, mu[e]);
#pragma unroll
      for (int e = 0; e < PS; e++) __stcs(h + (n + blk * PS + e) * kTile, S[e]);
      h -= NE * kTile;
      until = ck;
    }
    until--;
  }
  // status = 1 + index of the first block (in (time, block) order) that failed, as BdFam::step reports it
  int best_t = t_fail, best_blk = blk;
#pragma unroll
  for (int k = 0; k < nb; k++) {
    const int tk = __shfl_sync(kFull, t_fail, k * 16 + i);
    if (tk < best_t || (tk == best_t && k < best_blk)) {
      best_t = tk;
      best_blk = k;
    }
  }
  if (valid && blk == 0 && a.status) a.status[b] = best_t == 0x7fffffff ? 0 : best_blk + 1;
}

template <class Fam, int IG>
cudaError_t launch_forward_tpb(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  if (!a.perm16) return cudaErrorInvalidValue;
  const typename Fam::PriorT P = make_prior<Fam>(h);
  const long n_warps = (a.B + 255) / 256 * 16;
  const unsigned grid = (unsigned)((n_warps + 3) / 4);
  const BatchedPrior none{nullptr, nullptr, nullptr, nullptr};
  if (blocks_uniform<Fam>(P)) kalman_forward_tpb<Fam, IG, 0><<<grid, 128, 0, s>>>(P, a, none);
  else kalman_forward_tpb<Fam, IG, 1><<<grid, 128, 0, s>>>(P, a, none);
  return cudaGetLastError();
}

// per-system priors (block-diagonal for every system: checked by the caller), probde interrogation
template <class Fam>
cudaError_t launch_forward_tpb_pp(const HostPrior &h, const BatchedPrior &bq, const SolveArgs &a, cudaStream_t s) {
  if (!a.perm16) return cudaErrorInvalidValue;
  const typename Fam::PriorT P = make_prior<Fam>(h);
  const long n_warps = (a.B + 255) / 256 * 16;
  kalman_forward_tpb<Fam, IG_PROBDE, 2><<<(unsigned)((n_warps + 3) / 4), 128, 0, s>>>(P, a, bq);
  return cudaGetLastError();
}

template <class Fam, int CK, bool DO_MEAN, int DO_VAR, bool DO_SIM>
cudaError_t launch_backward_tpb_mode(const typename Fam::PriorT &P, int pm, const BatchedPrior *bq, const SolveArgs &a,
                                     cudaStream_t s) {
  using L = TpbLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>;
  auto kern = pm == 0 ? kalman_backward_tpb<Fam, CK, 0, DO_MEAN, DO_VAR, DO_SIM>
              : pm == 1 ? kalman_backward_tpb<Fam, CK, 1, DO_MEAN, DO_VAR, DO_SIM>
                        : kalman_backward_tpb<Fam, CK, 2, DO_MEAN, DO_VAR, DO_SIM>;
  // measurement knob: extra dynamic shared memory per CTA lowers the resident CTAs per SM without a rebuild
  static const size_t pad = [] {
    const char *v = getenv("RODEO_TPB_SMEM_PAD");
    return v ? (size_t)atol(v) : (size_t)0;
  }();
  const size_t smem = L::smem_bytes + pad;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  GranuleMaps maps;
  constexpr int n = Fam::n;
  if ((e = encode_granule_map(&maps.mu, DO_MEAN ? a.mu_out : nullptr, n, a.n_steps, a.B)) != cudaSuccess) return e;
  if ((e = encode_granule_map(&maps.x, DO_SIM ? a.x_out : nullptr, n, a.n_steps, a.B)) != cudaSuccess) return e;
  if ((e = encode_granule_map(&maps.var, DO_VAR ? a.var_out : nullptr, L::WV, a.n_steps, a.B, L::SVar::kG)) != cudaSuccess) return e;
  const long n_warps = (a.B + 255) / 256 * 16;
  const unsigned grid = (unsigned)((n_warps + L::WARPS - 1) / L::WARPS);
  const BatchedPrior none{nullptr, nullptr, nullptr, nullptr};
  kern<<<grid, L::THREADS, smem, s>>>(P, a, maps, bq ? *bq : none);
  return cudaGetLastError();
}

// Needs SolveArgs::perm16 = 1 (the forward pass wrote the history in the permuted order).
template <class Fam, int CK, bool DO_MEAN, int DO_VAR, bool DO_SIM>
cudaError_t launch_backward_tpb(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  using L = TpbLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>;
  if (!a.perm16) return cudaErrorInvalidValue;
  if ((double)16 * a.n_steps * Fam::n * Fam::n * 8 >= 2147483648.0) return cudaErrorInvalidValue;   // 32-bit row coordinates
  const typename Fam::PriorT P = make_prior<Fam>(h);
  return launch_backward_tpb_mode<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>(P, blocks_uniform<Fam>(P) ? 0 : 1, nullptr, a, s);
}

// log-likelihood smoother (DRAW: of the draw, else of the mean): FP64-bound, no shared memory, full occupancy
template <class Fam, int CK, bool DRAW>
cudaError_t launch_backward_tpb_ll(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  using L = TpbLayout<Fam, CK, !DRAW, 0, DRAW>;
  if (!a.perm16) return cudaErrorInvalidValue;
  const typename Fam::PriorT P = make_prior<Fam>(h);
  const long n_warps = (a.B + 255) / 256 * 16;
  const unsigned grid = (unsigned)((n_warps + L::WARPS - 1) / L::WARPS);
  GranuleMaps maps;
  memset(&maps, 0, sizeof maps);
  const BatchedPrior none{nullptr, nullptr, nullptr, nullptr};
  if (blocks_uniform<Fam>(P)) kalman_backward_tpb<Fam, CK, 0, !DRAW, 0, DRAW, true><<<grid, L::THREADS, 0, s>>>(P, a, maps, none);
  else kalman_backward_tpb<Fam, CK, 1, !DRAW, 0, DRAW, true><<<grid, L::THREADS, 0, s>>>(P, a, maps, none);
  return cudaGetLastError();
}

template <class Fam, int CK, bool DO_MEAN, int DO_VAR, bool DO_SIM>
cudaError_t launch_backward_tpb_pp(const HostPrior &h, const BatchedPrior &bq, const SolveArgs &a, cudaStream_t s) {
  if (!a.perm16) return cudaErrorInvalidValue;
  if ((double)16 * a.n_steps * Fam::n * Fam::n * 8 >= 2147483648.0) return cudaErrorInvalidValue;
  return launch_backward_tpb_mode<Fam, CK, DO_MEAN, DO_VAR, DO_SIM>(make_prior<Fam>(h), 2, &bq, a, s);
}

}  // namespace rodeo
