// kernel_launch.cuh -- __global__ wrappers and launchers of the thread-per-system bodies.
#pragma once
#include "registry.cuh"

namespace rodeo {

constexpr int kThreadsPerBlock = 64;

template <int n, int m>
inline void fill_prior(Prior<n, m> &P, const HostPrior &h) {
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
}

template <int p, int nb>
inline void fill_prior(PriorBD<p, nb> &P, const HostPrior &h) {
  const int n = p * nb;
  for (int b = 0; b < nb; b++) {
    for (int i = 0; i < p; i++) {
      P.w[b][i] = h.W[b + (size_t)(b * p + i) * nb];
      for (int j = 0; j < p; j++) {
        P.Q[b][i * p + j] = h.Q[(b * p + i) + (size_t)(b * p + j) * n];
        if (i <= j) P.R[b][sidx<p>(i, j)] = h.R[(b * p + i) + (size_t)(b * p + j) * n];
      }
    }
  }
  for (int i = 0; i < n; i++) P.c[i] = h.c[i];
  P.tmin = h.tmin;
  P.tmax = h.tmax;
  P.n_eval = h.n_eval;
}

template <class Fam>
inline typename Fam::PriorT make_prior(const HostPrior &h) {
  typename Fam::PriorT P;
  fill_prior(P, h);
  return P;
}

}  // namespace rodeo
#include "kalman_staged.cuh"
#include "kalman_bdw.cuh"
#include "kalman_sc.cuh"
#include "kalman_tpb.cuh"
#include "kalman_tpx.cuh"
#include "kalman_wpb.cuh"
namespace rodeo {

template <class Fam, bool UNROLL, int IG>
__global__ void __launch_bounds__(kThreadsPerBlock)
kalman_forward(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;   // history position (coalesced stores)
  const long b = a.perm16 ? perm16(u) : u;
  if (b < a.B) Bodies<UNROLL>::template forward<Fam, IG>(P, a, b);
}

template <class Fam, bool UNROLL, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
__global__ void __launch_bounds__(kThreadsPerBlock)
kalman_backward(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) Bodies<UNROLL>::template backward<Fam, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(P, a, b);
}

// ---- per-system priors: every system brings its own Q, c, R, W ------------------
// BatchedPrior (registry.cuh) holds the device arrays.  The thread builds its private copy of the
// prior (registers / local memory instead of the constant bank) and runs the same bodies.

template <class Fam, bool UNROLL, int IG>
__global__ void __launch_bounds__(kThreadsPerBlock)
kalman_forward_pp(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ BatchedPrior bp,
                  const __grid_constant__ SolveArgs a) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  typename Fam::PriorT Pl = P;
  load_batched_prior(Pl, bp, b);
  Bodies<UNROLL>::template forward<Fam, IG>(Pl, a, b);
}

template <class Fam, bool UNROLL, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
__global__ void __launch_bounds__(kThreadsPerBlock)
kalman_backward_pp(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ BatchedPrior bp,
                   const __grid_constant__ SolveArgs a) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  typename Fam::PriorT Pl = P;
  load_batched_prior(Pl, bp, b);
  Bodies<UNROLL>::template backward<Fam, 1, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>(Pl, a, b);
}

inline dim3 grid_for(long B) { return dim3((unsigned)((B + kThreadsPerBlock - 1) / kThreadsPerBlock)); }

template <class Fam, bool UNROLL, int IG>
cudaError_t launch_forward(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  kalman_forward<Fam, UNROLL, IG><<<grid_for(a.perm16 ? (a.B + 255) / 256 * 256 : a.B), kThreadsPerBlock, 0, s>>>(
      make_prior<Fam>(h), a);
  return cudaGetLastError();
}

// Forward pass of the block-diagonal register families: one warp per diagonal block (kalman_wpb.cuh);
// RODEO_CUDA_FORWARD=legacy restores the thread-per-block (two blocks) / thread-per-system (three or more) kernels.
template <class Fam, int IG, bool TWO>
cudaError_t launch_forward_bd(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  // Measured (B200): two blocks, FN 65,536 x 400: 0.368 -> 0.337 ms.  Three or more blocks: a gain only where the
  // pass is bound by its history stores (full history: MSEIR 32,768 x 1000 1.99 -> 1.85 ms); with a checkpointed
  // history the per-step CTA barrier costs more than the extra warps bring (MSEIR interval 2: 0.95 -> 1.28 ms,
  // Lorenz63 1.87 -> 1.90 ms), so those keep one thread per system.
  const bool wpb = !prefer_legacy_forward() && (TWO || a.ck == 1 || forward_preference() == 2);
  if (wpb) return launch_forward_wpb<Fam, IG>(h, a, s);
  if constexpr (TWO) return launch_forward_tpb<Fam, IG>(h, a, s);
  else return launch_forward<Fam, true, IG>(h, a, s);
}

template <class Fam, bool UNROLL, int CK, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
cudaError_t launch_backward(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  kalman_backward<Fam, UNROLL, CK, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>
      <<<grid_for(a.B), kThreadsPerBlock, 0, s>>>(make_prior<Fam>(h), a);
  return cudaGetLastError();
}

template <class Fam, bool UNROLL, int IG>
cudaError_t launch_forward_pp(const HostPrior &h, const BatchedPrior &bp, const SolveArgs &a, cudaStream_t s) {
  kalman_forward_pp<Fam, UNROLL, IG><<<grid_for(a.B), kThreadsPerBlock, 0, s>>>(make_prior<Fam>(h), bp, a);
  return cudaGetLastError();
}
template <class Fam, bool UNROLL, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
cudaError_t launch_backward_pp(const HostPrior &h, const BatchedPrior &bp, const SolveArgs &a, cudaStream_t s) {
  kalman_backward_pp<Fam, UNROLL, DO_MEAN, DO_VAR, DO_SIM, LOGLIK>
      <<<grid_for(a.B), kThreadsPerBlock, 0, s>>>(make_prior<Fam>(h), bp, a);
  return cudaGetLastError();
}

// backward launchers of one family for checkpoint interval CK (slot CKI of the table)
template <class Fam, class Ode, bool BD, bool UNROLL, int CK, int CKI>
void fill_backward(KernelSet &k) {
#ifndef RODEO_TPX_U
#define RODEO_TPX_U 1      // systems per lane of the draw-only thread-per-block smoother (kalman_tpx.cuh)
#endif
#ifndef RODEO_BDW_MIN_BLOCKS
#define RODEO_BDW_MIN_BLOCKS 3
#endif
  if constexpr (UNROLL && BD && Ode::n_meas >= RODEO_BDW_MIN_BLOCKS) {
    // block-diagonal with >= 3 blocks: one warp per diagonal block, CTA per tile (measured:
    // 2.1x on MSEIR, 1.9x on Lorenz63 over thread-per-system; 8 % slower at 2 blocks).
    // Full history only (checkpoint interval 1): see kalman_bdw.cuh.
    // Draw-only smoothing (solve_sim): one thread per block, two systems per lane, checkpoint interval 1 or 2
    // (kalman_tpx.cuh).  Intervals > 1 exist for that entry point only; the API picks the interval per entry point.
    if constexpr (CK == 1) {
      if constexpr (!Fam::kUT) {   // HBM-bound: shared with the structured twin (registry.cu)
        k.backward[CKI][BK_MV] = launch_backward_bdw<Fam, true, true, false>;
        k.backward[CKI][BK_BOTH] = launch_backward_bdw<Fam, true, true, true>;
        k.backward_alt[CKI][BK_SIM] = launch_backward_bdw<Fam, false, false, true>;
        k.backward_vl[0][0] = launch_backward_bdw<Fam, true, true, false, 1>;   // compact variance layouts
        k.backward_vl[0][1] = launch_backward_bdw<Fam, true, true, true, 1>;
        k.backward_vl[1][0] = launch_backward_bdw<Fam, true, true, false, 2>;
        k.backward_vl[1][1] = launch_backward_bdw<Fam, true, true, true, 2>;
        k.vl_ck = 1;
        k.vl_perm = false;
      }
      k.backward[CKI][BK_SIM] = launch_backward_tpx<Fam, 1, RODEO_TPX_U>;
    } else if constexpr (CK == 2) {
      k.backward[CKI][BK_SIM] = launch_backward_tpx<Fam, 2, RODEO_TPX_U>;
      return;   // the other entry points of this family keep the full history
    } else {
      return;   // leaves this slot empty: the API falls back to interval 1 for this family
    }
  } else if constexpr (UNROLL && BD && Ode::n_meas == 2) {
    // two diagonal blocks (FitzHugh-Nagumo): one thread per block, history prefetched into registers,
    // outputs as whole 256-byte-aligned granules through TMA tensor stores (kalman_tpb.cuh); these
    // launchers need the history in the permuted order (backward_perm -> SolveArgs::perm16)
    // The smoothers that write per-step outputs are HBM-bound: the structured arithmetic buys them nothing and
    // measured 3 % slower (FN solve_mv 2.03 -> 2.10 ms: same register cap, one spill), so the structured sets
    // reuse the general family's instantiation there (bit-identical results; also halves the build).  The
    // FP64-bound kernels (forward filters, log-likelihood smoothers) do use the specialisation.
    // (the structured sets leave these slots empty here: registry.cu points them at the general set's launchers)
    if constexpr (!Fam::kUT) {
      k.backward[CKI][BK_MV] = launch_backward_tpb<Fam, CK, true, true, false>;
      k.backward[CKI][BK_SIM] = launch_backward_tpb<Fam, CK, false, false, true>;
      k.backward[CKI][BK_BOTH] = launch_backward_tpb<Fam, CK, true, true, true>;
      k.backward_perm[CKI][BK_MV] = k.backward_perm[CKI][BK_SIM] = k.backward_perm[CKI][BK_BOTH] = true;
      if constexpr (CK == 2) {   // compact variance layouts: diagonal blocks only / marginal variances only
        k.backward_vl[0][0] = launch_backward_tpb<Fam, 2, true, 2, false>;
        k.backward_vl[0][1] = launch_backward_tpb<Fam, 2, true, 2, true>;
        k.backward_vl[1][0] = launch_backward_tpb<Fam, 2, true, 3, false>;
        k.backward_vl[1][1] = launch_backward_tpb<Fam, 2, true, 3, true>;
        k.vl_ck = 2;
        k.vl_perm = true;
      }
      // the round-1 smoother, kept for A/B measurements (RODEO_CUDA_BACKWARD=staged)
      k.backward_alt[CKI][BK_MV] = launch_backward_staged<Fam, CK, true, true, false>;
      k.backward_alt[CKI][BK_SIM] = launch_backward_staged<Fam, CK, false, false, true>;
      k.backward_alt[CKI][BK_BOTH] = launch_backward_staged<Fam, CK, true, true, true>;
    }
  } else if constexpr (UNROLL) {  // registers: history and outputs staged through shared memory
    k.backward[CKI][BK_MV] = launch_backward_staged<Fam, CK, true, true, false>;
    k.backward[CKI][BK_SIM] = launch_backward_staged<Fam, CK, false, false, true>;
    k.backward[CKI][BK_BOTH] = launch_backward_staged<Fam, CK, true, true, true>;
  } else {
    k.backward[CKI][BK_MV] = launch_backward<Fam, UNROLL, CK, true, true, false, false>;
    k.backward[CKI][BK_SIM] = launch_backward<Fam, UNROLL, CK, false, false, true, false>;
    k.backward[CKI][BK_BOTH] = launch_backward<Fam, UNROLL, CK, true, true, true, false>;
  }
  if constexpr (UNROLL && BD && Ode::n_meas == 2 && CK == 2) {
    // two blocks, the auto interval: thread-per-block log-likelihood smoother (kalman_tpb.cuh), structured
    // arithmetic where the set has it (the kernel is FP64-bound); same sums, bit for bit, as the staged kernels
    k.backward[CKI][BK_LL_MEAN] = launch_backward_tpb_ll<Fam, CK, false>;
    k.backward[CKI][BK_LL_DRAW] = launch_backward_tpb_ll<Fam, CK, true>;
    k.backward_perm[CKI][BK_LL_MEAN] = k.backward_perm[CKI][BK_LL_DRAW] = true;
    k.backward_alt[CKI][BK_LL_MEAN] = launch_backward_staged<Fam, CK, true, false, false, true>;
    k.backward_alt[CKI][BK_LL_DRAW] = launch_backward_staged<Fam, CK, false, false, true, true>;
  } else if constexpr (UNROLL) {   // register families: history through the TMA ring, nothing staged out
    k.backward[CKI][BK_LL_MEAN] = launch_backward_staged<Fam, CK, true, false, false, true>;
    k.backward[CKI][BK_LL_DRAW] = launch_backward_staged<Fam, CK, false, false, true, true>;
  } else {
    k.backward[CKI][BK_LL_MEAN] = launch_backward<Fam, UNROLL, CK, true, false, false, true>;
    k.backward[CKI][BK_LL_DRAW] = launch_backward<Fam, UNROLL, CK, false, false, true, true>;
  }
}

// K = n_state for the dense family, block size p for the block-diagonal family.
template <class Ode, int K, bool BD, bool UNROLL, bool ST = false>
KernelSet make_kernel_set(const char *name) {
  using Fam = typename FamSel<UNROLL, BD, Ode, K, ST>::type;
  KernelSet k;
  k.structured = ST;
  k.mixed = Ode::mixed;
  static_assert(!(Ode::mixed && BD), "unequal block sizes run on the dense family");
  for (int a = 0; a < 8; a++) k.off[a] = a < Ode::n_meas ? Ode::template var_off<Fam::n>(a) : 0;
  k.ode = Ode::id;
  k.n = Fam::n;
  k.m = Ode::n_meas;
  k.blockdiag = BD;
  k.registers = UNROLL;
  k.name = name;
  k.forward[IG_PROBDE] = launch_forward<Fam, UNROLL, IG_PROBDE>;
  k.forward[IG_SCHOBERT] = launch_forward<Fam, UNROLL, IG_SCHOBERT>;
  k.forward[IG_CHKREBTII] = launch_forward<Fam, UNROLL, IG_CHKREBTII>;
  k.forward_perm[0] = k.forward_perm[1] = k.forward_perm[2] = nullptr;
  if constexpr (UNROLL && BD && Ode::n_meas == 2) {   // warp-per-block forward (kalman_wpb.cuh), permuted order
    k.forward_perm[IG_PROBDE] = launch_forward_bd<Fam, IG_PROBDE, true>;
    k.forward_perm[IG_SCHOBERT] = launch_forward_bd<Fam, IG_SCHOBERT, true>;
    k.forward_perm[IG_CHKREBTII] = launch_forward_bd<Fam, IG_CHKREBTII, true>;
  }
  if constexpr (UNROLL && BD && Ode::n_meas >= 3) {   // warp-per-block forward, plain order
    k.forward[IG_PROBDE] = launch_forward_bd<Fam, IG_PROBDE, false>;
    k.forward[IG_SCHOBERT] = launch_forward_bd<Fam, IG_SCHOBERT, false>;
    k.forward[IG_CHKREBTII] = launch_forward_bd<Fam, IG_CHKREBTII, false>;
  }
  for (int c = 0; c < kNumCk; c++)
    for (int b = 0; b < BK_COUNT; b++) {
      k.backward[c][b] = k.backward_alt[c][b] = nullptr;
      k.backward_perm[c][b] = false;
    }
  k.backward_vl[0][0] = k.backward_vl[0][1] = k.backward_vl[1][0] = k.backward_vl[1][1] = nullptr;
  k.vl_ck = 1;
  k.vl_perm = false;
  fill_backward<Fam, Ode, BD, UNROLL, 1, 0>(k);
  if constexpr (UNROLL) {  // checkpointed history for the register families only
    fill_backward<Fam, Ode, BD, UNROLL, 2, 1>(k);
    fill_backward<Fam, Ode, BD, UNROLL, 4, 2>(k);
  }
  k.hist_elems = Fam::NE;
  k.sc_tables = nullptr;
  k.sc_table_doubles = nullptr;
  k.sc_table_var_offset = nullptr;
  k.sc_forward = nullptr;
  for (int b = 0; b < BK_COUNT; b++) k.sc_backward[b] = nullptr;
  k.pp_forward = nullptr;
  for (int b = 0; b < BK_COUNT; b++) k.pp_backward[b] = nullptr;
  k.pp_bd_forward = nullptr;
  for (int b = 0; b < BK_COUNT; b++) k.pp_bd_backward[b] = nullptr;
  if constexpr (UNROLL && BD && Ode::n_meas == 2 && !ST) {   // per-system priors that are block diagonal per system
    k.pp_bd_forward = launch_forward_tpb_pp<Fam>;             // (general arithmetic; structured twins borrow them)
    k.pp_bd_backward[BK_MV] = launch_backward_tpb_pp<Fam, 2, true, true, false>;
    k.pp_bd_backward[BK_SIM] = launch_backward_tpb_pp<Fam, 2, false, false, true>;
    k.pp_bd_backward[BK_BOTH] = launch_backward_tpb_pp<Fam, 2, true, true, true>;
  }
  k.wide_forward[0] = k.wide_forward[1] = k.wide_forward[2] = nullptr;
  for (int b = 0; b < BK_COUNT; b++) k.wide_backward[b] = nullptr;
  k.wide_pp_forward = nullptr;
  for (int b = 0; b < BK_COUNT; b++) k.wide_pp_backward[b] = nullptr;
  k.wide_hist_elems = 0;
  if constexpr (!BD) {  // per-system priors: dense family only (a batched prior need not be structured)
    k.pp_forward = launch_forward_pp<Fam, UNROLL, IG_PROBDE>;
    if constexpr (UNROLL) {   // register family: the staged kernel with a thread-private prior (coalesced / TMA stores)
      k.pp_backward[BK_MV] = launch_backward_staged_pp<Fam, true, true, false>;
      k.pp_backward[BK_SIM] = launch_backward_staged_pp<Fam, false, false, true>;
      k.pp_backward[BK_BOTH] = launch_backward_staged_pp<Fam, true, true, true>;
    } else {
      k.pp_backward[BK_MV] = launch_backward_pp<Fam, UNROLL, true, true, false, false>;
      k.pp_backward[BK_SIM] = launch_backward_pp<Fam, UNROLL, false, false, true, false>;
      k.pp_backward[BK_BOTH] = launch_backward_pp<Fam, UNROLL, true, true, true, false>;
    }
    k.pp_backward[BK_LL_MEAN] = launch_backward_pp<Fam, UNROLL, true, false, false, true>;
    k.pp_backward[BK_LL_DRAW] = launch_backward_pp<Fam, UNROLL, false, false, true, true>;
  }
  if constexpr (UNROLL) {  // shared-covariance mode for the register families
    k.sc_tables = launch_sc_tables<Fam>;
    k.sc_table_doubles = sc_table_doubles<Fam>;
    k.sc_table_var_offset = sc_table_var_offset<Fam>;
    k.sc_forward = launch_sc_forward<Fam>;
    k.sc_backward[BK_MV] = launch_sc_backward<Fam, true, false, false>;
    k.sc_backward[BK_SIM] = launch_sc_backward<Fam, false, true, false>;
    k.sc_backward[BK_BOTH] = launch_sc_backward<Fam, true, true, false>;
    k.sc_backward[BK_LL_MEAN] = launch_sc_backward<Fam, true, false, true>;
    k.sc_backward[BK_LL_DRAW] = launch_sc_backward<Fam, false, true, true>;
  }
  return k;
}

}  // namespace rodeo
