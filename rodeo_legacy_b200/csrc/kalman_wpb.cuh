// kalman_wpb.cuh -- forward filter of the block-diagonal family with one WARP PER DIAGONAL BLOCK of a 32-system
// history tile: CTA = n_blocks warps, warp k runs block k of the tile's 32 systems, lane = position in the tile.
//
// Why (measured on B200, see NOTES_r02.md):
//   * one thread per SYSTEM leaves the machine empty for the long, narrow batches: Lorenz63, 16,384 systems = 512
//     warps = 3.5 per SM, each running its three blocks one after the other (1.87 ms for 8.9 GB of history);
//   * one thread per block inside a warp (kalman_tpb.cuh: lane = blk * 16 + i) triples / doubles the warps but its
//     history stores are pieces of a row (128 bytes for 2 blocks, 80 bytes for 3), and every lane evaluates the whole
//     right-hand side although it needs one component (the branch on blk would diverge inside the warp).
// Here the block index is WARP-uniform: the right-hand side is specialised per component without divergence
// (FitzHugh-Nagumo: one division instead of two per lane and step), every history store is a whole 256-byte row
// of the tile, and there are n_blocks times the warps of the thread-per-system kernel.  The blocks of a system
// meet once per step: each warp publishes the first component of its predicted mean (or of the chkrebtii draw) in
// shared memory, one CTA barrier, every warp reads the n_blocks values its component of the right-hand side needs.
// The slots are double buffered by step parity, so one barrier per step is enough.
//
// Same routines as BdFam::step (predict_mean / predict_var / update_block), so the filtered moments are
// bit-identical to the other forward kernels'.  Handles both history orders (SolveArgs::perm16).
#pragma once
#include "kalman_tpb.cuh"

namespace rodeo {

// component k of the right-hand side, k warp-uniform: the full functor is inlined per case and the unused
// components are dead code (same expressions, same contraction: bit-identical to eval()[k])
template <class Ode, int n>
__device__ __forceinline__ double ode_component(int k, const double *X, double t, const double *th) {
  double y[Ode::n_meas];
  double out = 0.0;
#pragma unroll
  for (int c = 0; c < Ode::n_meas; c++) {
    if (k == c) {
      Ode::template eval<n>(X, t, th, y);
      out = y[c];
    }
  }
  return out;
}

template <class Fam, int IG, int PM>
__global__ void __launch_bounds__(Fam::nb * 32)
kalman_forward_wpb(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ SolveArgs a) {
  using Ode = typename Fam::Ode;
  constexpr int p = Fam::p, nb = Fam::nb, n = Fam::n, PS = Fam::PS, NE = Fam::NE;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  __shared__ double lead_s[2][nb][32];
  __shared__ int fail_s[nb][32];
  const int lane = threadIdx.x % 32, blk = threadIdx.x / 32;   // blk is warp-uniform
  const long u = (long)blockIdx.x * 32 + lane;                 // position in the history
  const long b = a.perm16 ? perm16(u) : u;                     // the system stored there
  const bool valid = b < a.B;
  const long bs = valid ? b : 0;                               // a safe system for the loads of idle lanes
  BlockPrior<Fam, PM == 0> bp(P, blk);
  const int N = P.n_eval, ck = a.ck, R = n_records(N, ck);
  double mu[p], S[PS], th[NT];
#pragma unroll
  for (int e = 0; e < p; e++) mu[e] = a.x0[bs * n + blk * p + e];
#pragma unroll
  for (int e = 0; e < PS; e++) S[e] = 0.0;
#pragma unroll
  for (int e = 0; e < NT; e++) th[e] = (Ode::n_theta > 0) ? a.theta[bs * Ode::n_theta + e] : 0.0;
  const double *zb = (IG == IG_CHKREBTII) ? a.z + bs * a.z_stride + blk * p : nullptr;
  // records are stored from the last one (r = R - 1, the earliest time index kept) down to r = 0 (time index N):
  // a running pointer, no division by the run-time interval inside the loop
  double *h = a.hist + hist_offset(u, R - 1, R, NE);
  int t_fail = 0x7fffffff;     // first time index at which this block's innovation variance was not positive
  int back = N - 1;            // distance of time index t + 1 from the end
  int until = back % ck;       // steps until the next stored record
  // shared-state addresses of the exchange slots, computed once (left to the compiler, the thread index is
  // re-read with S2R in every step: ncu, 6 % of the stall samples)
  const uint32_t lead_wr = smem_u32(&lead_s[0][blk][lane]), lead_rd = smem_u32(&lead_s[0][0][lane]);
  uint32_t par = 0;            // byte offset of the slots of this step's parity
  auto publish = [&](double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(lead_wr + par), "d"(v) : "memory"); };
  for (int t = 0; t < N; t++, back--) {
    double mup[p], Sp[PS];
    reg::predict_mean<p, Fam::kUT>(bp.Q(), bp.c(), mu, mup);
    int f = 0;
    if (IG != IG_CHKREBTII) publish(mup[0]);   // before the covariance work
    reg::predict_var<p, false, Fam::kUT>(bp.Q(), bp.R(), S, Sp, nullptr);
    if (IG == IG_CHKREBTII) {   // x_state ~ N(mu_pred, Sigma_pred) with z[:, t]  (KalmanODE.pyx:13-49)
      double xs[p], V[PS], zt[p];
#pragma unroll
      for (int e = 0; e < PS; e++) V[e] = Sp[e];
#pragma unroll
      for (int e = 0; e < p; e++) zt[e] = zb[(long)t * n + e];
      f = reg::state_sim<p>(xs, mup, V, zt);
      publish(xs[0]);
    }
    // the gain needs only Sigma_pred: its division runs while the blocks meet (update_block = gain + apply)
    double A[p], inv;
    const int f2 = reg::update_block_gain<p, Fam::kWK>(bp.w(), Sp, IG == IG_SCHOBERT, A, &inv);
    __syncthreads();
    double X[n];
#pragma unroll
    for (int e = 0; e < n; e++) X[e] = 0.0;
#pragma unroll
    for (int k = 0; k < nb; k++)
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(X[k * p]) : "r"(lead_rd + par + (uint32_t)(k * 32 * 8)) : "memory");
    par ^= (uint32_t)(nb * 32 * 8);
    const double tt = P.tmin + (P.tmax - P.tmin) * (t + 1) / P.n_eval;
    const double ya = ode_component<Ode, n>(blk, X, tt, th);
    reg::update_block_apply<p, Fam::kWK>(bp.w(), mup, Sp, ya, A, inv);
    if ((f | f2) && t_fail == 0x7fffffff) t_fail = t;
#pragma unroll
    for (int e = 0; e < p; e++) mu[e] = mup[e];
#pragma unroll
    for (int e = 0; e < PS; e++) S[e] = Sp[e];
    if (until == 0) {
      // (idle lanes of the last tile store the safe system's record: the padded positions hold finite numbers)
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
  fail_s[blk][lane] = t_fail;
  __syncthreads();
  if (blk == 0 && valid && a.status) {
    int best_t = 0x7fffffff, best_blk = 0;
#pragma unroll
    for (int k = 0; k < nb; k++) {
      const int tk = fail_s[k][lane];
      if (tk < best_t) {
        best_t = tk;
        best_blk = k;
      }
    }
    a.status[b] = best_t == 0x7fffffff ? 0 : best_blk + 1;
  }
}

// RODEO_CUDA_FORWARD=legacy keeps the previous forward kernels, =wpb forces this one (A/B measurements)
inline int forward_preference() {
  static const int v = [] {
    const char *e = getenv("RODEO_CUDA_FORWARD");
    return !e ? 0 : e[0] == 'l' ? 1 : e[0] == 'w' ? 2 : 0;
  }();
  return v;
}
inline bool prefer_legacy_forward() { return forward_preference() == 1; }

template <class Fam, int IG>
cudaError_t launch_forward_wpb(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  const typename Fam::PriorT P = make_prior<Fam>(h);
  const long positions = a.perm16 ? (a.B + 255) / 256 * 256 : a.B;
  const unsigned grid = (unsigned)((positions + 31) / 32);
  if (blocks_uniform<Fam>(P)) kalman_forward_wpb<Fam, IG, 0><<<grid, Fam::nb * 32, 0, s>>>(P, a);
  else kalman_forward_wpb<Fam, IG, 1><<<grid, Fam::nb * 32, 0, s>>>(P, a);
  return cudaGetLastError();
}

}  // namespace rodeo
