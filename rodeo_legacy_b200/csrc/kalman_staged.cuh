// kalman_staged.cuh -- backward (smoother) kernel of the register family with the
// filter history staged through shared memory by bulk async copies (TMA, cp.async.bulk +
// mbarrier) and the outputs transposed through shared memory: TMA tensor stores for n_state = 6,
// coalesced 16-byte stores otherwise; the log-likelihood variant emits nothing per step.
//
// One warp owns one tile of 32 systems (one thread per system) for the whole backward pass:
//   * its history record for one (checkpointed) time index is NE*32 contiguous doubles in HBM
//     (layout in kalman_common.cuh).  Lane 0 issues ONE cp.async.bulk per record into a two-slot
//     ring in shared memory, two records ahead of the one being processed, so the
//     DRAM latency of the read stream is hidden without spending registers on prefetch;
//     each lane then reads its own column of the record (bank-conflict free).
//   * every lane writes the step's outputs (mean, covariance, draw) into its own row of a
//     staging tile; the tile then leaves as one cp.async.bulk.tensor per array (TmaEmitter: the
//     tensor map {W, T1, B} with box {W, 1, 32} describes exactly the 32 runs of a step), or the
//     warp stores it so that each system's contiguous run in the user-visible layout
//     out[b][t][...]  is written by neighbouring lanes as 16-byte pieces (StagedEmitter).
//   * BlockedRuns (used by the shared-covariance and warp-per-block kernels) collects K steps of a short
//     per-step run per system and stores them as one sector-aligned piece.
//   * with checkpointing (CK > 1) the filter is re-run from each checkpoint record into CK-1
//     scratch records in shared memory (same layout, each lane touches only its own column),
//     which the smoother then consumes exactly like loaded records.
// Warps never synchronise with each other (no __syncthreads after the prologue).
#pragma once
#include <cuda.h>   // CUtensorMap and its enums only: the encoder is fetched through the runtime, no libcuda link
#include <cstdint>
#include <cstring>
#include <type_traits>

#include "kalman_tps.cuh"
#include "registry.cuh"

namespace rodeo {

// ---- TMA tensor stores for the outputs (n_state = 6) ---------------------------------------------
// out[b][t][0..W) viewed as a 3-D tensor {W, T1, B}; one box {W, 1, 32} = the W-double runs of the 32
// systems of a tile at one time index, i.e. exactly what a warp emits per step.  The staging tile in
// shared memory is the box image ([32][W] doubles, dense), so ONE cp.async.bulk.tensor per array and
// step replaces the tile -> register -> global flush (36 LDS.128 + 36 STG.128 per warp-step for FN
// solve_mv) and the hardware clips the systems beyond B.  Needs 16-byte multiples everywhere:
// W * 8 % 16 == 0, i.e. even n_state.
struct StoreMaps {
  alignas(64) CUtensorMap mu, var, x;
};

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline encode_tiled_fn tensor_map_encoder() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  }
  return fn;
}

// Tensor map of one output array: W doubles per (system, time index).
inline cudaError_t encode_out_map(CUtensorMap *m, double *base, int W, long T1, long B) {
  if (!base) {
    memset(m, 0, sizeof *m);
    return cudaSuccess;
  }
  encode_tiled_fn enc = tensor_map_encoder();
  if (!enc) return cudaErrorNotSupported;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)T1, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 8, (cuuint64_t)W * 8 * (cuuint64_t)T1};
  const cuuint32_t box[3] = {(cuuint32_t)W, 1, (cuuint32_t)kTile};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// shared -> global tensor store of one box (bulk async-group completion)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(static_cast<uint32_t>(__cvta_generic_to_shared(src_smem))), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every committed store has finished READING shared memory (the tile may be overwritten)
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// every committed store is complete
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Orders this thread's earlier generic-proxy accesses to shared memory (LDS of a ring slot) before
// later async-proxy accesses (the TMA refill of that slot).  Without it an LDS that is still
// queued in the memory pipe when the refill lands returns the NEW record: observed on B200 as
// one 32-system tile per ~10^4 picking up record r+2's tail elements (found by the
// "covariance is bit-identical across the batch" property test).
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

#ifndef RODEO_STAGED_WARPS
#define RODEO_STAGED_WARPS 2
#endif
constexpr int kStagedWarps = RODEO_STAGED_WARPS;  // warps per CTA
constexpr int kStagedThreads = kStagedWarps * 32;

template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK = false>
struct StagedLayout {
  static constexpr int n = Fam::n, NE = Fam::NE;
  static constexpr int REC = NE * kTile;                                         // doubles per record
  static constexpr int OUTW = (DO_MEAN ? n : 0) + (DO_VAR ? n * n : 0) + (DO_SIM ? n : 0);
  // The staging tile can be emitted in one pass or in two (pass 0 = mean, draw and the first H
  // variance entries, pass 1 = the rest), which halves the tile (FN solve_mv: 10.5 -> 6.5 KB per
  // warp) and lets 5-7 CTAs fit per SM instead of 4.  MEASURED on B200 (FN solve_mv, checkpoint 2):
  // one pass 2.32 ms; two passes split mid-sector 3.07-3.34 ms; two passes split on a 32-byte sector
  // boundary with the register cap that realises the extra occupancy 2.50 ms.  The second
  // write/sync/read round costs more than the occupancy buys: one pass it is.
  static constexpr bool TWO_PASS = false;
  // variance entries in pass 0: about half, rounded up to whole 32-byte sectors of the run so that
  // no sector is written by both passes
  static constexpr int H = TWO_PASS ? ((n * n / 2 + 3) / 4 * 4) : n * n;
  static constexpr int VEC_W = (DO_MEAN ? n : 0) + (DO_SIM ? n : 0);
  static constexpr int W0 = VEC_W + (DO_VAR ? H : 0), W1 = TWO_PASS ? n * n - H : 0;
  static constexpr int ROW = ((W0 > W1 ? W0 : W1) + 1) / 2 * 2;                  // row stride (even)
  // per warp: 2 ring slots (checkpoint records, TMA), CK-1 scratch records (re-run filter),
  // the output staging tile, 2 mbarriers
  // outputs leave through TMA tensor stores (dense per-array tiles) when every run is a 16-byte multiple and
  // the dense rows cost at most a 2-way bank conflict on the 16-byte row writes: n_state = 6
  static constexpr bool kTma = (n == 6) && !TWO_PASS && !LOGLIK;
  // per-warp region kept a multiple of 128 bytes (TMA source alignment)
  static constexpr int STAGE = LOGLIK ? 0 : kTma ? kTile * (VEC_W + (DO_VAR ? n * n : 0)) : kTile * ROW;   // staging tile(s)
  static constexpr int WARP_DOUBLES = ((2 + (CK - 1)) * REC + STAGE + 2 + 15) / 16 * 16;
  static constexpr size_t smem_bytes = (size_t)kStagedWarps * WARP_DOUBLES * sizeof(double);
  // No min-blocks launch bound: with a one-pass tile shared memory limits the kernel to 4 CTAs per
  // SM anyway, and a register cap below what ptxas wants only adds spills (tried 168 / 128:
  // 2.5 -> 5.1 ms).
  static constexpr int MINB = 1;
};

// Generic (any parity) store of columns [C0, C0+LEN) of every system's run of W doubles from a
// staging tile (system s at tile[s*TW + OFF ..]) to out + ((b0+s)*T1 + t)*W + C0, by NTHREADS threads.
template <int W, int C0, int LEN, int OFF, int TW, int NTHREADS>
__device__ __forceinline__ void flush_part(const double *tile, double *out, long b0, long B, long T1, int t,
                                           int tid) {
  constexpr bool V2 = (W % 2 == 0) && (C0 % 2 == 0) && (LEN % 2 == 0) && (OFF % 2 == 0) && (TW % 2 == 0);
  if (V2) {
    constexpr int C = LEN / 2;
#pragma unroll
    for (int q0 = 0; q0 < kTile * C; q0 += NTHREADS) {
      const int q = q0 + tid;
      if (q < kTile * C) {
        const int s = q / C, c = q - s * C;
        if (b0 + s < B) {
          const double2 v = *reinterpret_cast<const double2 *>(tile + s * TW + OFF + 2 * c);
          *reinterpret_cast<double2 *>(out + ((b0 + s) * T1 + t) * W + C0 + 2 * c) = v;
        }
      }
    }
  } else {
#pragma unroll
    for (int q0 = 0; q0 < kTile * LEN; q0 += NTHREADS) {
      const int q = q0 + tid;
      if (q < kTile * LEN) {
        const int s = q / LEN, c = q - s * LEN;
        if (b0 + s < B) out[((b0 + s) * T1 + t) * W + C0 + c] = tile[s * TW + OFF + c];
      }
    }
  }
}

// Per-lane constants of the tile -> global store of columns [C0, C0+LEN) of one output array
// (run of W doubles per system, staged at stage[s*ROW + OFF ..], destination
// out + ((b0+s)*T1 + t)*W + C0).  The piece is cut into 16-byte chunks, C = LEN/2 per system;
// SPI = 32/C systems are stored per instruction, so a lane always handles the same chunk c of the
// same relative system s: every address is lane-constant + t*W + j*const, computed with one
// multiply-add per emit instead of a divide / multiply chain per chunk (the integer work of the
// generic flush was half of all instructions of the kernel).
template <int W, int C0, int LEN, int OFF, int ROW>
struct FlushPlan {
  static constexpr bool kFast = (W % 2 == 0) && (C0 % 2 == 0) && (LEN % 2 == 0) && (OFF % 2 == 0) &&
                                (ROW % 2 == 0) && (LEN / 2 <= 32) && (LEN > 0);
  static constexpr int C = kFast ? LEN / 2 : 1;
  static constexpr int SPI = kFast ? 32 / C : 1;          // systems per store instruction
  static constexpr int NI = (kTile + SPI - 1) / SPI;      // store instructions per emit
  const double *src;   // stage + s_l*ROW + OFF + 2*c_l
  double *dst0;        // out + ((b0+s_l)*T1)*W + C0 + 2*c_l   (time index 0)
  long jstride;        // SPI*T1*W doubles between consecutive instructions
  int s_l, rem;        // relative system of this lane; valid systems in this tile
  bool active, full;
  __device__ __forceinline__ void init(const double *stage, double *out, long b0, long B, long T1, int lane) {
    if constexpr (kFast) {
      s_l = lane / C;
      const int c_l = lane - s_l * C;
      active = lane < SPI * C;
      full = b0 + kTile <= B;
      src = stage + s_l * ROW + OFF + 2 * c_l;
      dst0 = out ? out + ((b0 + s_l) * T1) * W + C0 + 2 * c_l : nullptr;
      jstride = (long)SPI * T1 * W;
      rem = (int)((B - b0 < kTile ? B - b0 : kTile));
    }
  }
  __device__ __forceinline__ void run(const double *stage, double *out, long b0, long B, long T1, int t,
                                      int lane) const {
    if constexpr (LEN == 0) {
      return;
    } else if constexpr (kFast) {
      double *d = dst0 + (long)t * W;
      if (active) {
        if (full && (kTile % SPI == 0)) {
#pragma unroll
          for (int j = 0; j < NI; j++)
            *reinterpret_cast<double2 *>(d + j * jstride) = *reinterpret_cast<const double2 *>(src + j * SPI * ROW);
        } else {
#pragma unroll
          for (int j = 0; j < NI; j++)
            if (j * SPI + s_l < rem)
              *reinterpret_cast<double2 *>(d + j * jstride) = *reinterpret_cast<const double2 *>(src + j * SPI * ROW);
        }
      }
    } else {
      flush_part<W, C0, LEN, OFF, ROW, 32>(stage, out, b0, B, T1, t, lane);
    }
  }
};

// Short per-step runs (a mean or a draw: W = n doubles, 48-120 bytes) written one per visit straddle
// 32-byte sectors and leave the memory system at less than half speed (tools/microbench/write_pattern.cu:
// 48-byte runs 2.2 TB/s, the same bytes as sector-aligned multi-step pieces 4.7-4.9 TB/s).  BlockedRuns
// collects K consecutive time indices of a system in a shared-memory row and stores them as ONE piece of
// K*W doubles.  Blocks are cut on the GLOBAL run index g = b*T1 + t (block = K runs with g / K equal), so
// every full block starts at a multiple of K*W*8 bytes whatever the parity of T1 and b: systems of one
// tile complete their blocks at different steps, a few rows per step.  Trajectory ends give partial blocks.
// Rows are written by their owner thread (put), stored cooperatively by NT threads after a barrier (store).
template <int W, int K, int NT>
struct BlockedRuns {
  static_assert((K & (K - 1)) == 0 && (K * W) % 2 == 0 && kTile % K == 0, "K a power of two, K*W even");
  static constexpr int RS = K * W, PR = RS / 2;     // doubles and 16-byte pieces per row
  static constexpr int DOUBLES = kTile * RS;
  // steady state with T1 odd: the K residue classes of the row index complete their blocks in turn, one class
  // (kTile / K rows) per step; thread tid handles piece q = j*NT + tid of the class, q = i*PR + p
  static constexpr int CLS_PIECES = (kTile / K) * PR, JN = (CLS_PIECES + NT - 1) / NT;
  double *tile, *out;
  long b0, B, T1;
  int N, r0;                 // r0 = (b0*T1) mod K
  bool odd;
  __device__ __forceinline__ void init(double *tile_, double *out_, long b0_, long B_, long T1_, int) {
    tile = tile_;
    out = out_;
    b0 = b0_;
    B = B_;
    T1 = T1_;
    N = (int)T1_ - 1;
    r0 = (int)((b0_ * T1_) & (K - 1));
    odd = (T1_ & 1) != 0;
  }
  __device__ __forceinline__ void put(int s, int t, const double *v) const {
    const int slot = (int)(((b0 + s) * T1 + t) & (K - 1));
    double *d = tile + s * RS + slot * W;
    if constexpr (W % 2 == 0) {
#pragma unroll
      for (int i = 0; i < W / 2; i++) reinterpret_cast<double2 *>(d)[i] = make_double2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < W; i++) d[i] = v[i];
    }
  }
  // part of a run (entries off .. off+CNT-1), e.g. one diagonal block's share of the state vector
  template <int CNT>
  __device__ __forceinline__ void put_part(int s, int t, int off, const double *v) const {
    const int slot = (int)(((b0 + s) * T1 + t) & (K - 1));
    double *d = tile + s * RS + slot * W + off;
#pragma unroll
    for (int i = 0; i < CNT; i++) d[i] = v[i];
  }
  // any T1, trajectory ends: every piece looks up its row's state
  __device__ __noinline__ void store_general(int t, int tid) const {
    for (int q = tid; q < kTile * PR; q += NT) {
      const int s = q / PR, p = q - s * PR;
      const long g = (b0 + s) * T1 + t;
      const int slot = (int)(g & (K - 1));
      if (b0 + s >= B || (slot != 0 && t != 0)) continue;     // this row's block is not complete yet
      int hi = slot + (N - t);                                // last valid slot (the block may run past time N)
      hi = hi < K - 1 ? hi : K - 1;
      const int d0 = 2 * p, k0 = d0 / W, k1 = (d0 + 1) / W;
      const bool v0 = k0 >= slot && k0 <= hi, v1 = k1 >= slot && k1 <= hi;
      const double *src = tile + s * RS + d0;
      double *dst = out + (g - slot) * W + d0;
      if (v0 && v1) *reinterpret_cast<double2 *>(dst) = *reinterpret_cast<const double2 *>(src);
      else if (v0) dst[0] = src[0];
      else if (v1) dst[1] = src[1];
    }
  }
  __device__ __forceinline__ void store(int t, int tid) const {
    if (!odd || t == 0 || N - t < K - 1) {
      store_general(t, tid);
      return;
    }
    // the class c with ((b0 + c)*T1 + t) mod K == 0; T1 odd, so T1*T1 = 1 (mod 8): c = -(r0 + t)*T1 (mod K)
    const int cls = (int)((-(long)(r0 + t) * T1) & (K - 1));
    const double *src = tile + cls * RS;
    double *dst = out + ((b0 + cls) * T1 + t) * W;
    const int rows_left = (int)(B - b0 - cls);                // row b0 + cls + i*K exists while i*K < rows_left
    const int strideW = (int)T1 * W;
#pragma unroll
    for (int j = 0; j < JN; j++) {
      const int q = j * NT + tid, i = q / PR, p = q - i * PR;  // piece p of the class's i-th row (few integer ops,
      if (q < CLS_PIECES && i * K < rows_left)                 // cheaper than holding per-thread offset tables)
        *reinterpret_cast<double2 *>(dst + (long)(i * K) * strideW + 2 * p) =
            *reinterpret_cast<const double2 *>(src + i * K * RS + 2 * p);
    }
  }
};

template <class Fam, class L, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
struct StagedEmitter {
  static constexpr int n = Fam::n, ROW = L::ROW, H = L::H;
  static constexpr int OFF_X = DO_MEAN ? n : 0, OFF_VAR = L::VEC_W;
  FlushPlan<n, 0, (DO_MEAN ? n : 0), 0, ROW> pm;
  FlushPlan<n, 0, (DO_SIM ? n : 0), OFF_X, ROW> px;
  FlushPlan<n * n, 0, (DO_VAR ? H : 0), OFF_VAR, ROW> pv0;
  FlushPlan<n * n, H, L::W1, 0, ROW> pv1;
  double *stage;
  long b0, T1;
  int lane;
  __device__ __forceinline__ void init(double *stage_, const SolveArgs &a, long b0_, long T1_, int lane_) {
    stage = stage_;
    b0 = b0_;
    T1 = T1_;
    lane = lane_;
    if (DO_MEAN) pm.init(stage, a.mu_out, b0, a.B, T1, lane);
    if (DO_SIM) px.init(stage, a.x_out, b0, a.B, T1, lane);
    if (DO_VAR) pv0.init(stage, a.var_out, b0, a.B, T1, lane);
    if (L::TWO_PASS) pv1.init(stage, a.var_out, b0, a.B, T1, lane);
  }
  __device__ __forceinline__ void put_row(const double *v) const {
    double2 *row = reinterpret_cast<double2 *>(stage + lane * ROW);
#pragma unroll
    for (int i = 0; i < ROW / 2; i++) row[i] = make_double2(v[2 * i], v[2 * i + 1]);
  }
  // Emit time index t of the whole tile: every lane writes its row (16-byte shared stores), the
  // warp stores the tile; long rows go out in two passes through the same tile.
  __device__ __forceinline__ void emit(const SolveArgs &a, int t, const double *mus, const double *Ss,
                                       const double *x) const {
    {
      double v[ROW];
#pragma unroll
      for (int i = 0; i < ROW; i++) v[i] = 0.0;
      if (DO_MEAN) {
#pragma unroll
        for (int i = 0; i < n; i++) v[i] = mus[i];
      }
      if (DO_SIM) {
#pragma unroll
        for (int i = 0; i < n; i++) v[OFF_X + i] = x[i];
      }
      if (DO_VAR) {
#pragma unroll
        for (int e = 0; e < H; e++) v[OFF_VAR + e] = Fam::var_at(Ss, e % n, e / n);
      }
      put_row(v);
    }
    __syncwarp();
    if (DO_MEAN) pm.run(stage, a.mu_out, b0, a.B, T1, t, lane);
    if (DO_SIM) px.run(stage, a.x_out, b0, a.B, T1, t, lane);
    if (DO_VAR) pv0.run(stage, a.var_out, b0, a.B, T1, t, lane);
    __syncwarp();
    if (L::TWO_PASS) {
      double v[ROW];
#pragma unroll
      for (int i = 0; i < ROW; i++) v[i] = 0.0;
#pragma unroll
      for (int e = H; e < n * n; e++) v[e - H] = Fam::var_at(Ss, e % n, e / n);
      put_row(v);
      __syncwarp();
      pv1.run(stage, a.var_out, b0, a.B, T1, t, lane);
      __syncwarp();
    }
  }
};

// Emission through TMA tensor stores: per-array dense tiles [32][W] (mean, draw, variance, each starting
// on a 128-byte boundary); every lane writes its rows with 16-byte shared stores, lane 0 issues one tensor
// store per array.  The previous step's stores must have finished reading the tiles before they are
// overwritten (wait_group.read).  Measured alternatives (FN solve_mv backward kernel, 2.35 ms with the LSU
// flush): all arrays through the TMA 2.25 ms; variance through the TMA and the 48-byte mean runs through the
// LSU flush 2.26 ms; mean runs paired into sector-aligned 96-byte pieces (registers hold the previous step)
// 2.32 ms.  At 2.25 ms the kernel moves its 10.7 GB at 4.77 TB/s, which is what a no-compute kernel
// writing the same 288-byte-run pattern reaches (tools/microbench/write_pattern.cu: 4.74 TB/s).
template <class Fam, class L, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
struct TmaEmitter {
  static constexpr int n = Fam::n;
  static constexpr int OFF_MU = 0, OFF_X = OFF_MU + (DO_MEAN ? kTile * n : 0), OFF_VAR = OFF_X + (DO_SIM ? kTile * n : 0);
  static_assert((kTile * n) % 16 == 0, "tiles must start on 128-byte boundaries");
  double *stage;
  const StoreMaps *maps;
  int b0, lane;
  __device__ __forceinline__ void init(double *stage_, const StoreMaps *maps_, long b0_, int lane_) {
    stage = stage_;
    maps = maps_;
    b0 = (int)b0_;
    lane = lane_;
  }
  template <int W>
  __device__ __forceinline__ void put(double *tile, const double *v) const {
    double2 *row = reinterpret_cast<double2 *>(tile + lane * W);
#pragma unroll
    for (int i = 0; i < W / 2; i++) row[i] = make_double2(v[2 * i], v[2 * i + 1]);
  }
  __device__ __forceinline__ void emit(const SolveArgs &, int t, const double *mus, const double *Ss,
                                       const double *x) const {
    if (lane == 0) tma_wait_read();
    __syncwarp();
    if (DO_MEAN) put<n>(stage + OFF_MU, mus);
    if (DO_SIM) put<n>(stage + OFF_X, x);
    if (DO_VAR) {
      double2 *row = reinterpret_cast<double2 *>(stage + OFF_VAR + lane * n * n);
#pragma unroll
      for (int i = 0; i < n * n / 2; i++)
        row[i] = make_double2(Fam::var_at(Ss, (2 * i) % n, (2 * i) / n), Fam::var_at(Ss, (2 * i + 1) % n, (2 * i + 1) / n));
    }
    fence_proxy_async();   // the rows (generic proxy) become visible to the TMA (async proxy) ...
    __syncwarp();          // ... of every lane, before lane 0 issues the stores
    if (lane == 0) {
      if (DO_MEAN) tma_store_3d(&maps->mu, stage + OFF_MU, 0, t, b0);
      if (DO_SIM) tma_store_3d(&maps->x, stage + OFF_X, 0, t, b0);
      if (DO_VAR) tma_store_3d(&maps->var, stage + OFF_VAR, 0, t, b0);
      tma_commit();
    }
  }
  __device__ __forceinline__ void finish() const {
    if (lane == 0) tma_wait_all();
    __syncwarp();
  }
};

// Per-system priors: device arrays, batch-major, each system laid out like the constructor's arrays
// (column-major); nullptr = keep the handle's shared value.
template <int n, int m>
__device__ __forceinline__ void load_batched_prior(Prior<n, m> &P, const BatchedPrior &bp, long b) {
  if (bp.Q) {
    const double *q = bp.Q + b * n * n;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) P.Q[i * n + j] = q[i + j * n];
  }
  if (bp.R) {
    const double *r = bp.R + b * n * n;
    for (int i = 0; i < n; i++)
      for (int j = i; j < n; j++) P.R[sidx<n>(i, j)] = r[i + j * n];
  }
  if (bp.c) {
    for (int i = 0; i < n; i++) P.c[i] = bp.c[b * n + i];
  }
  if (bp.W) {
    const double *w = bp.W + b * m * n;
    for (int a = 0; a < m; a++)
      for (int j = 0; j < n; j++) P.W[a * n + j] = w[a + j * m];
  }
}

// No per-step output at all: the thinned Gaussian log-likelihood of the smoothed mean (or of the draw) is
// accumulated per system (examples/inference.py:54-56, 66-94) and one double per system leaves the kernel.
template <class Fam, bool DO_SIM>
struct LoglikEmitter {
  double lp;
  int d;
  __device__ __forceinline__ void init(const SolveArgs &a) {
    lp = 0.0;
    d = a.n_obs_times - 1;
  }
  __device__ __forceinline__ void emit(const SolveArgs &a, int t, const double *mus, const double *, const double *x) {
    while (d >= 0 && a.thin_index[d] == t) {
      lp += reg::loglik_terms<Fam::n>(a, d, DO_SIM ? x : mus);
      d--;
    }
  }
  __device__ __forceinline__ void finish(const SolveArgs &a, long b, bool valid) const {
    // constant of the normal density: -(log sqrt(2 pi) + log gamma) per observation
    const double cst = -0.91893853320467274178 - log(a.gamma);
    if (valid) a.loglik[b] = lp + cst * (double)(a.n_obs_times * a.n_obs_comp);
  }
};

// The whole backward pass of one warp-tile.  P is the kernel's __grid_constant__ prior (every coefficient an
// immediate constant-bank operand after inlining) or, for per-system priors, a thread-private copy.
template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
__device__ __forceinline__ void staged_backward_body(const typename Fam::PriorT &P, const SolveArgs &a,
                                                     const StoreMaps &maps) {
  using L = StagedLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>;
  using Ode = typename Fam::Ode;
  constexpr int n = Fam::n, NS = Fam::NS, NE = Fam::NE, REC = L::REC;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  extern __shared__ __align__(128) double smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long tile = (long)blockIdx.x * kStagedWarps + warp;
  const long n_tiles = (a.B + kTile - 1) / kTile;
  if (tile >= n_tiles) return;  // warp-uniform; warps are independent from here on
  double *ring = smem + (size_t)warp * L::WARP_DOUBLES;
  double *scratch = ring + 2 * REC;
  double *stage = scratch + (CK - 1) * REC;
  uint64_t *bar = reinterpret_cast<uint64_t *>(stage + L::STAGE);
  const int N = P.n_eval, R = n_records(N, CK);
  const long T1 = (long)N + 1, b0 = tile * kTile, b = b0 + lane;
  const bool valid = b < a.B;
  const long bs = valid ? b : b0;  // a safe system index for loads of inactive lanes
  const double *hist_tile = a.hist + (size_t)tile * R * REC;
  constexpr uint32_t kRecBytes = REC * sizeof(double);

#ifdef RODEO_DEBUG_POISON
  for (int i = lane; i < 2 * REC; i += 32) ring[i] = __longlong_as_double(0x7ff8000000000000LL);
  __syncwarp();
#endif
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
  typename std::conditional<
      LOGLIK, LoglikEmitter<Fam, DO_SIM>,
      typename std::conditional<L::kTma, TmaEmitter<Fam, L, DO_MEAN, DO_VAR, DO_SIM>,
                                StagedEmitter<Fam, L, DO_MEAN, DO_VAR, DO_SIM>>::type>::type em;
  if constexpr (LOGLIK) em.init(a);
  else if constexpr (L::kTma) em.init(stage, &maps, b0, lane);
  else em.init(stage, a, b0, T1, lane);
  uint32_t phase = 0;
  // wait for record r (ring slot r & 1); returns this lane's column of it
  auto acquire = [&](int r) -> const double * {
    const int k = r & 1;
    mbar_wait(&bar[k], (phase >> k) & 1u);
    phase ^= 1u << k;
#ifdef RODEO_DEBUG_RING
    if (r == 0) {  // does the ring slot hold exactly what HBM holds for record r? (last element only)
      const double sm = *(const volatile double *)(ring + k * REC + lane + (NE - 1) * kTile);
      const double g = *(const volatile double *)(hist_tile + (size_t)r * REC + lane + (NE - 1) * kTile);
      if (g != sm && valid && a.status) atomicMax(a.status + b, 100000 + r * 100 + (NE - 1));
    }
#endif
    return ring + k * REC + lane;
  };
  // every lane is done with record r: refill its slot with record r + 2
  auto release = [&](int r) {
    fence_proxy_async();   // all of this lane's reads of the slot are complete ...
    __syncwarp();          // ... for every lane, before lane 0 lets the TMA overwrite it
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
    em.emit(a, N, mus, Ss, x);
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
      em.emit(a, base, mus, Ss, x);
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
          em.emit(a, t, mus, Ss, x);
        }
      }
      if (base >= 1) {
        const int f = Fam::template bw_step<kTile, DO_MEAN, DO_VAR, DO_SIM>(
            P, rec, DO_SIM ? zb + (long)(base - 1) * n : nullptr, mus, Ss, x);
        if (f && !fail) fail = f;
        release(r);
        em.emit(a, base, mus, Ss, x);
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
    em.emit(a, 0, mus, Ss, x);
  }
  if constexpr (LOGLIK) em.finish(a, b, valid);
  else if constexpr (L::kTma) em.finish();
  if (valid && a.status && fail && a.status[b] == 0) a.status[b] = fail;
}

template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK = false>
__global__ void __launch_bounds__(kStagedThreads, StagedLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>::MINB)
kalman_backward_staged(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a,
                       const __grid_constant__ StoreMaps maps) {
  staged_backward_body<Fam, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(P, a, maps);
}

// Per-system priors (dense family, full history): every thread overlays its system's Q, c, R, W on a private
// copy of the shared prior and runs the same body -- same TMA ring, same staged / TMA-store emission.
template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
__global__ void __launch_bounds__(kStagedThreads)
kalman_backward_staged_pp(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ BatchedPrior bp,
                          const __grid_constant__ SolveArgs a, const __grid_constant__ StoreMaps maps) {
  const long b0 = ((long)blockIdx.x * kStagedWarps + threadIdx.x / 32) * kTile, b = b0 + threadIdx.x % 32;
  typename Fam::PriorT Pl = P;
  if (b0 < a.B) load_batched_prior(Pl, bp, b < a.B ? b : b0);
  staged_backward_body<Fam, 1, DO_MEAN, DO_VAR, DO_SIM, false>(Pl, a, maps);
}

template <class Fam, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK = false>
cudaError_t launch_backward_staged(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  using L = StagedLayout<Fam, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>;
  auto kern = kalman_backward_staged<Fam, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem_bytes);
  if (e != cudaSuccess) return e;
  if (a.perm16) return cudaErrorInvalidValue;   // the permuted history order belongs to kalman_tpb.cuh
  const long n_tiles = (a.B + kTile - 1) / kTile;
  const unsigned grid = (unsigned)((n_tiles + kStagedWarps - 1) / kStagedWarps);
  StoreMaps maps;
  memset(&maps, 0, sizeof maps);
  if (L::kTma) {
    constexpr int n = Fam::n;
    if (DO_MEAN && (e = encode_out_map(&maps.mu, a.mu_out, n, a.n_steps, a.B)) != cudaSuccess) return e;
    if (DO_SIM && (e = encode_out_map(&maps.x, a.x_out, n, a.n_steps, a.B)) != cudaSuccess) return e;
    if (DO_VAR && (e = encode_out_map(&maps.var, a.var_out, n * n, a.n_steps, a.B)) != cudaSuccess) return e;
  }
  kern<<<grid, kStagedThreads, L::smem_bytes, s>>>(make_prior<Fam>(h), a, maps);
  return cudaGetLastError();
}

template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM>
cudaError_t launch_backward_staged_pp(const HostPrior &h, const BatchedPrior &bp, const SolveArgs &a, cudaStream_t s) {
  using L = StagedLayout<Fam, 1, DO_MEAN, DO_VAR, DO_SIM, false>;
  auto kern = kalman_backward_staged_pp<Fam, DO_MEAN, DO_VAR, DO_SIM>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem_bytes);
  if (e != cudaSuccess) return e;
  const long n_tiles = (a.B + kTile - 1) / kTile;
  const unsigned grid = (unsigned)((n_tiles + kStagedWarps - 1) / kStagedWarps);
  StoreMaps maps;
  memset(&maps, 0, sizeof maps);
  if (L::kTma) {
    constexpr int n = Fam::n;
    if (DO_MEAN && (e = encode_out_map(&maps.mu, a.mu_out, n, a.n_steps, a.B)) != cudaSuccess) return e;
    if (DO_SIM && (e = encode_out_map(&maps.x, a.x_out, n, a.n_steps, a.B)) != cudaSuccess) return e;
    if (DO_VAR && (e = encode_out_map(&maps.var, a.var_out, n * n, a.n_steps, a.B)) != cudaSuccess) return e;
  }
  typename Fam::PriorT P;
  fill_prior(P, h);
  kern<<<grid, kStagedThreads, L::smem_bytes, s>>>(P, bp, a, maps);
  return cudaGetLastError();
}

}  // namespace rodeo
