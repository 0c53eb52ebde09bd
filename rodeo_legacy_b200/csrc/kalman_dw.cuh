// kalman_dw.cuh -- dense WIDE-state kernels (7 <= n_state <= 16, any Q, R, W): one 16-lane
// group per system, lane j owns COLUMN j of every n x n matrix.
//
// A 15 x 15 covariance does not fit one thread's registers (the thread-per-system register
// family stops at n = 6) and the rolled local-memory fallback is ~100x off.  Here a system's
// matrices are spread over the lanes of half a warp:
//   * lane j holds column j of Sigma_f, Sigma_p (then its Cholesky factor), T = Q Sigma_f (then
//     T~), D / Sigma_s -- n doubles each -- and element j of the mean vectors; full mean vectors
//     are replicated in every lane (cheap, avoids exchanges);
//   * products with the constant prior matrices (Q, W: immediate constant-bank operands) are
//     purely local:  (Q X)[:, j] = Q X[:, j];
//   * products of two data matrices go through a per-system shared-memory scratch: the owner lanes
//     write their columns, everyone reads what it needs (broadcast reads, conflict free):
//     T^T (to form Sigma_p = Q T^T + R), the Cholesky factor (triangular solves), D and T~;
//   * the column Cholesky broadcasts each finished column with warp shuffles (width 16);
//   * history record per step: n rows of Sigma_f + one row of mu_f, each row = 32 doubles (two
//     systems x 16 lanes), coalesced both ways; outputs are staged through shared memory so a
//     system's (n*n)-run is written by consecutive lanes.
// Two systems per warp; lanes >= n of a group idle (n = 15: 6 %, n = 9: 44 %).  This family is the
// GENERAL path for wide states; block-diagonal priors use the much cheaper BdFam kernels.
//
// Arithmetic: rodeo/eigen/KalmanTV.h predict :280-294, update :324-357, smooth_mv :447-467,
// smooth_sim :503-533, state_sim :619-628; probde / schobert interrogation
// (rodeo/eigen/KalmanTVODE.h:333-341).
#pragma once
#include "registry.cuh"

namespace rodeo {

constexpr int kDwLanes = 16;     // lanes per system
constexpr int kDwWarps = 4;      // warps per CTA (8 systems)

template <class Ode_, int n_>
struct DwFam {
  using Ode = Ode_;
  static constexpr int n = n_, m = Ode::n_meas;
  static constexpr int ROWS = n_ + 1;                 // history rows per record: Sigma_f rows + mu_f
  static constexpr int NE = ROWS * kDwLanes;          // history doubles per system per record
  static constexpr int LD = (n_ % 2) ? n_ : n_ + 1;   // odd leading dimension of the scratch matrices
  static constexpr int SCR = 3 * n_ * LD + n_ * n_;   // per system: XA, XB, XC (n x LD) + OUT (n*n)
  static constexpr int CONST = 2 * n_ * n_ + Ode_::n_meas * n_ + n_;   // R, W, Q, c (per CTA, or per system)
  using PriorT = Prior<n_, Ode_::n_meas>;
};

__device__ __forceinline__ double dw_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src, kDwLanes); }
__device__ __forceinline__ double dw_sum(double v) {
#pragma unroll
  for (int o = kDwLanes / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, kDwLanes);
  return v;
}

// Element j of c + Q x for a replicated vector x: row j of Q and c[j] are lane-indexed, so they come
// from the shared copies (Qs row-major n x n, cs).  Computing the n elements on n lanes and
// all-gathering them costs n FMA + 2n shuffles per lane instead of n*n replicated FMA.
template <int n>
__device__ __forceinline__ double dw_affine_row(const double *Qs, const double *cs, const double *x, int j) {
  if (j >= n) return 0.0;
  double s = cs[j];
#pragma unroll
  for (int k = 0; k < n; k++) s += Qs[j * n + k] * x[k];
  return s;
}

// out[:, j] = Q in[:, j]   (column owned by this lane)
template <int n>
__device__ __forceinline__ void dw_qcol(const double *Q, const double *in, double *out) {
#pragma unroll
  for (int i = 0; i < n; i++) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < n; k++) s += Q[i * n + k] * in[k];
    out[i] = s;
  }
}

// P[:, j] = Q (T^T)[:, j] + R[:, j] = Q * (row j of T) + R[:, j].  T is distributed by columns: the
// transpose goes through the scratch matrix X (n x LD).  Rs = shared copy of R (row-major n x n).
template <int n, int LD>
__device__ __forceinline__ void dw_predict_cov(const double *Q, const double *Rs, const double *Tcol, double *X,
                                               int j, bool act, double *Pcol) {
  if (act) {
#pragma unroll
    for (int i = 0; i < n; i++) X[i * LD + j] = Tcol[i];
  }
  __syncwarp();
  double Trow[n];
#pragma unroll
  for (int k = 0; k < n; k++) Trow[k] = act ? X[j * LD + k] : 0.0;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < n; i++) {
    double s = act ? Rs[i * n + j] : 0.0;
#pragma unroll
    for (int k = 0; k < n; k++) s += Q[i * n + k] * Trow[k];
    Pcol[i] = s;
  }
}

// Column Cholesky of a symmetric matrix distributed by columns (lane j: A[:, j]); on exit lane j
// holds column j of the lower factor L (entries i >= j; entries i < j are garbage) and
// invd[k] = 1 / L[k][k] replicated.  Inactive lanes carry an identity column.
template <int n>
__device__ __forceinline__ int dw_chol(double *A, double *invd, int j) {
  int fail = 0;
#pragma unroll
  for (int k = 0; k < n; k++) {
    // the pivot column k is final in lane k: normalise it there, then broadcast it
    double d = dw_bcast(A[k], k);
    if (!(d > 0.0) && !fail) fail = k + 1;
    const double r = inv_sqrt(d);
    invd[k] = r;
    double colk[n];
#pragma unroll
    for (int i = k; i < n; i++) {
      const double lik = dw_bcast(A[i], k) * r;   // L[i][k]
      colk[i] = lik;
      if (j == k) A[i] = lik;
    }
    // trailing update of this lane's column j > k:  A[i][j] -= L[i][k] L[j][k]  (i >= j)
    double ljk = 0.0;
#pragma unroll
    for (int i = k + 1; i < n; i++) ljk = (i == j) ? colk[i] : ljk;
    if (j > k) {
#pragma unroll
      for (int i = k + 1; i < n; i++) A[i] -= colk[i] * ljk;
    }
  }
  return fail;
}

// x <- (L L^T)^{-1} x for this lane's column x (n entries); L read from the scratch matrix
// XL[i*LD + k] = L[i][k] (i >= k), invd replicated.
template <int n, int LD>
__device__ __forceinline__ void dw_solve(const double *XL, const double *invd, double *x) {
#pragma unroll
  for (int i = 0; i < n; i++) {
    double s = x[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= XL[i * LD + k] * x[k];
    x[i] = s * invd[i];
  }
#pragma unroll
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
#pragma unroll
    for (int k = i + 1; k < n; k++) s -= XL[k * LD + i] * x[k];
    x[i] = s * invd[i];
  }
}

// write this lane's column into a scratch matrix (X[i][j] = col[i])
template <int n, int LD>
__device__ __forceinline__ void dw_put(double *X, const double *col, int j, bool act) {
  if (act) {
#pragma unroll
    for (int i = 0; i < n; i++) X[i * LD + j] = col[i];
  }
}

// Shared copies of the lane-indexed prior constants: R (full, row-major), W, Q, c.  One copy per
// CTA, or -- per-system priors (PP) -- one per system, filled by that system's 16 lanes from the
// batched device arrays (column-major per system, nullptr = the handle's shared value).
template <class Fam, bool PP>
__device__ __forceinline__ void dw_load_consts(const typename Fam::PriorT &P, const BatchedPrior &bp, long b,
                                               double *dst) {
  constexpr int n = Fam::n, m = Fam::m;
  double *Rs = dst, *Ws = Rs + n * n, *Qs = Ws + m * n, *cs = Qs + n * n;
  const int first = PP ? (int)(threadIdx.x % kDwLanes) : (int)threadIdx.x;
  const int step = PP ? kDwLanes : (int)blockDim.x;
  for (int e = first; e < n * n; e += step) {
    const int i = e / n, k = e % n;
    Rs[e] = (PP && bp.R) ? bp.R[b * n * n + i + k * n] : P.R[sidx<n>(i, k)];
    Qs[e] = (PP && bp.Q) ? bp.Q[b * n * n + i + k * n] : P.Q[e];
  }
  for (int e = first; e < m * n; e += step) {
    const int q = e / n, k = e % n;
    Ws[e] = (PP && bp.W) ? bp.W[b * m * n + q + k * m] : P.W[e];
  }
  for (int e = first; e < n; e += step) cs[e] = (PP && bp.c) ? bp.c[b * n + e] : P.c[e];
  __syncthreads();
}

template <class Fam>
__device__ __forceinline__ long dw_hist_offset(long b, int r, int R) {
  // [pair][record][row][32]: pair = b / 2, lane slot = (b % 2) * 16 + j
  return (((b >> 1) * R + r) * (long)Fam::ROWS) * 32 + (b & 1) * kDwLanes;
}

template <class Fam, int IG, bool PP>
__global__ void __launch_bounds__(kDwWarps * 32)
kalman_dw_forward(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ BatchedPrior bp,
                  const __grid_constant__ SolveArgs a) {
  using Ode = typename Fam::Ode;
  constexpr int n = Fam::n, m = Fam::m, LD = Fam::LD;
  constexpr int NT = Ode::n_theta > 0 ? Ode::n_theta : 1;
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, lane = tid % 32, j = lane % kDwLanes;
  const int sys_in_cta = tid / kDwLanes;
  const long b = (long)blockIdx.x * (kDwWarps * 2) + sys_in_cta;
  const bool valid = b < a.B, act = valid && j < n;
  const long bs = valid ? b : 0;
  double *cst = PP ? smem + (size_t)sys_in_cta * (Fam::SCR + Fam::CONST) : smem;
  double *X = PP ? cst + Fam::CONST : smem + Fam::CONST + (size_t)sys_in_cta * Fam::SCR;   // this system's scratch
  double *XA = X, *XB = X + n * LD;                   // XB reused as the (m x n) image of A
  dw_load_consts<Fam, PP>(P, bp, bs, cst);
  const double *Rs = cst, *Ws = Rs + n * n, *Qs = Ws + m * n, *cs = Qs + n * n;
  const double *Qu = PP ? Qs : P.Q, *Wu = PP ? Ws : P.W;   // operands every lane reads at the same index
  double mu[n], Sf[n], th[NT];
#pragma unroll
  for (int i = 0; i < n; i++) {
    mu[i] = a.x0[bs * n + i];
    Sf[i] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < NT; i++) th[i] = (Ode::n_theta > 0) ? a.theta[bs * Ode::n_theta + i] : 0.0;
  const int N = P.n_eval;
  int fail = 0;
  for (int t = 0; t < N; t++) {
    double mup[n], T[n], Pc[n], y[m];
    const double mupj = dw_affine_row<n>(Qs, cs, mu, j);
#pragma unroll
    for (int i = 0; i < n; i++) mup[i] = dw_bcast(mupj, i);
    dw_qcol<n>(Qu, Sf, T);
    dw_predict_cov<n, LD>(Qu, Rs, T, XA, j, act, Pc);
    const double tt = P.tmin + (P.tmax - P.tmin) * (t + 1) / N;
    if (IG == IG_CHKREBTII) {   // x_state ~ N(mu_pred, Sigma_pred) with z[:, t]  (KalmanODE.pyx:13-49)
      double V[n], vinv[n], xs[n];
#pragma unroll
      for (int i = 0; i < n; i++) V[i] = act ? Pc[i] : ((i == j) ? 1.0 : 0.0);   // identity columns in idle lanes
      const int fc = dw_chol<n>(V, vinv, j);
      if (fc && !fail) fail = fc;
      // x = mu_p + L z: lane k contributes L[:, k] z[k]; the row sums are all-reduced, so every lane gets the vector
      const double zk = act ? a.z[bs * a.z_stride + (long)t * n + j] : 0.0;
#pragma unroll
      for (int i = 0; i < n; i++) xs[i] = mup[i] + dw_sum((act && i >= j) ? V[i] * zk : 0.0);
      Ode::template eval<n>(xs, tt, th, y);
    } else {
      Ode::template eval<n>(mup, tt, th, y);
    }
    // A[:, j] = W P[:, j]
    double Ac[m];
#pragma unroll
    for (int q = 0; q < m; q++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < n; k++) s += Wu[q * n + k] * Pc[k];
      Ac[q] = act ? s : 0.0;
    }
    // S = A W^T (+ H = A W^T under probde), reduced over the lanes
    double Sm[Sym<m>::size], invd[m];
#pragma unroll
    for (int q = 0; q < m; q++)
#pragma unroll
      for (int r = q; r < m; r++) {
        const double part = (j < n) ? Ac[q] * Ws[r * n + j] : 0.0;   // W[r][j]: lane-indexed, from shared
        const double s = dw_sum(part);
        Sm[sidx<m>(q, r)] = (IG == IG_SCHOBERT) ? s : s + s;
      }
    int f = reg::chol_packed<m>(Sm, invd);
    if (f && !fail) fail = f;
    double e[m], Kc[m];
#pragma unroll
    for (int q = 0; q < m; q++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < n; k++) s += Wu[q * n + k] * mup[k];
      e[q] = y[q] - s;
      Kc[q] = Ac[q];
    }
    reg::chol_solve_vec<m>(Sm, invd, Kc, 1);            // K~[:, j]
    double mufj = mupj;
#pragma unroll
    for (int q = 0; q < m; q++) mufj += Kc[q] * e[q];
    // Sigma_f[:, j] = P[:, j] - A^T K~[:, j]: needs every column of A -> scratch (m x n image)
    if (act) {
#pragma unroll
      for (int q = 0; q < m; q++) XB[q * LD + j] = Ac[q];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < n; i++) {
      double s = Pc[i];
#pragma unroll
      for (int q = 0; q < m; q++) s -= XB[q * LD + i] * Kc[q];
      Sf[i] = act ? s : 0.0;
    }
    __syncwarp();
    // new replicated mean
#pragma unroll
    for (int i = 0; i < n; i++) mu[i] = dw_bcast(mufj, i);
    // history record (time index t+1 = record N-1-t): rows of Sigma_f, then mu_f
    double *h = a.hist + dw_hist_offset<Fam>(valid ? b : bs, N - 1 - t, N) + j;
    if (valid) {
#pragma unroll
      for (int i = 0; i < n; i++) h[i * 32] = Sf[i];
      h[n * 32] = mufj;
    }
  }
  if (valid && j == 0 && a.status) a.status[b] = fail;
}

// Backward pass (smooth_mv / smooth_sim / both); with LOGLIK no per-step output is written: the thinned
// Gaussian log-likelihood of the mean (DO_MEAN) or the draw (DO_SIM) is accumulated instead, each lane
// handling the observed components it owns (examples/inference.py:54-56, 66-83).
template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK, bool PP>
__global__ void __launch_bounds__(kDwWarps * 32)
kalman_dw_backward(const __grid_constant__ typename Fam::PriorT P, const __grid_constant__ BatchedPrior bp,
                   const __grid_constant__ SolveArgs a) {
  constexpr int n = Fam::n, LD = Fam::LD;
  extern __shared__ __align__(128) double smem[];
  constexpr int m = Fam::m;
  const int tid = threadIdx.x, lane = tid % 32, j = lane % kDwLanes;
  const int sys_in_cta = tid / kDwLanes;
  const long b = (long)blockIdx.x * (kDwWarps * 2) + sys_in_cta;
  const bool valid = b < a.B, act = valid && j < n;
  const long bs = valid ? b : 0;
  double *cst = PP ? smem + (size_t)sys_in_cta * (Fam::SCR + Fam::CONST) : smem;
  double *X = PP ? cst + Fam::CONST : smem + Fam::CONST + (size_t)sys_in_cta * Fam::SCR;
  double *XA = X, *XB = X + n * LD, *XC = X + 2 * n * LD, *XO = X + 3 * n * LD;
  dw_load_consts<Fam, PP>(P, bp, bs, cst);
  const double *Rs = cst, *Qs = Rs + n * n + m * n, *cs = Qs + n * n;
  const double *Qu = PP ? Qs : P.Q;
  const int N = P.n_eval;
  const long T1 = (long)N + 1;
  const double *zb = DO_SIM ? a.z + bs * a.z_stride : nullptr;
  double *mu_o = (DO_MEAN && !LOGLIK) ? a.mu_out + bs * T1 * n : nullptr;
  double *var_o = (DO_VAR && !LOGLIK) ? a.var_out + bs * T1 * n * n : nullptr;
  double *x_o = (DO_SIM && !LOGLIK) ? a.x_out + bs * T1 * n : nullptr;
  double lpj = 0.0;                      // this lane's share of the log-likelihood
  int d = LOGLIK ? a.n_obs_times - 1 : -1;
  double musj = 0.0, xsj = 0.0, Ss[n];   // element j of the smoothed mean / draw; column j of Sigma_s
  int fail = 0;

  // Writes time index t.  The covariance column is symmetrised on the way out (and in place, so the
  // recursion carries the same bits it reported): column-wise arithmetic leaves Sigma(i, j) and
  // Sigma(j, i) a rounding apart, and every other kernel family returns exactly symmetric matrices.
  auto emit = [&](int t, double musj, double xsj, double *Scol) {
    if (LOGLIK) {
      while (d >= 0 && a.thin_index[d] == t) {
        for (int k = 0; k < a.n_obs_comp; k++) {
          if (a.state_index[k] == j) {
            const double r = (a.Y[d * a.n_obs_comp + k] - (DO_SIM ? xsj : musj)) / a.gamma;
            lpj -= 0.5 * r * r;
          }
        }
        d--;
      }
      return;
    }
    if (DO_MEAN && act) mu_o[(long)t * n + j] = musj;
    if (DO_SIM && act) x_o[(long)t * n + j] = xsj;
    if (DO_VAR) {
      if (act) {
#pragma unroll
        for (int i = 0; i < n; i++) XO[j * n + i] = Scol[i];   // XO[j][i] = Sigma(i, j)
      }
      __syncwarp();
      if (act) {
#pragma unroll
        for (int i = 0; i < n; i++) Scol[i] = 0.5 * (Scol[i] + XO[i * n + j]);
      }
      if (valid) {
        for (int e = j; e < n * n; e += kDwLanes) {
          const int r = e / n, c = e - r * n;
          var_o[(long)t * n * n + e] = 0.5 * (XO[e] + XO[c * n + r]);
        }
      }
      __syncwarp();
    }
  };

  {  // terminal time index N (record 0)
    const double *h = a.hist + dw_hist_offset<Fam>(bs, 0, N) + j;
    double Sf[n];
#pragma unroll
    for (int i = 0; i < n; i++) Sf[i] = act ? h[i * 32] : ((i == j) ? 1.0 : 0.0);
    const double mufj = act ? h[n * 32] : 0.0;
    if (DO_SIM) {
      double V[n], invd[n];
#pragma unroll
      for (int i = 0; i < n; i++) V[i] = Sf[i];
      int f = dw_chol<n>(V, invd, j);
      if (f && !fail) fail = f;
      // x = mu + L z: lane k contributes L[:, k] z[k]; row sums over the lanes
      const double zk = act ? zb[(long)(N - 1) * n + j] : 0.0;
      double xj = 0.0;
#pragma unroll
      for (int i = 0; i < n; i++) {
        const double s = dw_sum((act && i >= j) ? V[i] * zk : 0.0);
        xj = (i == j) ? s : xj;
      }
      xsj = xj + mufj;
    }
    musj = mufj;
#pragma unroll
    for (int i = 0; i < n; i++) Ss[i] = Sf[i];
    emit(N, musj, xsj, Ss);
  }
  for (int t = N - 1; t >= 1; t--) {
    const double *h = a.hist + dw_hist_offset<Fam>(bs, N - t, N) + j;
    double Sf[n], T[n], Pc[n], invd[n];
#pragma unroll
    for (int i = 0; i < n; i++) Sf[i] = act ? h[i * 32] : 0.0;
    const double mufj = act ? h[n * 32] : 0.0;
    double mupj;
    {
      double muf[n];
#pragma unroll
      for (int i = 0; i < n; i++) muf[i] = dw_bcast(mufj, i);
      mupj = dw_affine_row<n>(Qs, cs, muf, j);
    }
    dw_qcol<n>(Qu, Sf, T);                                   // T[:, j] = Q Sigma_f[:, j]
    dw_predict_cov<n, LD>(Qu, Rs, T, XA, j, act, Pc);        // Sigma_p[:, j]
    if (!act) {                                               // identity column keeps the Cholesky finite
#pragma unroll
      for (int i = 0; i < n; i++) Pc[i] = (i == j) ? 1.0 : 0.0;
    }
    if (DO_VAR && act) {   // D = Sigma_s+ - Sigma_p (symmetric) straight into its scratch matrix
#pragma unroll
      for (int i = 0; i < n; i++) XC[i * LD + j] = Ss[i] - Pc[i];
    }
    int f = dw_chol<n>(Pc, invd, j);                           // Pc now holds L[:, j]
    if (f && !fail) fail = f;
    // L into the scratch matrix XA (row-major L[i][k]) for the triangular solves
    if (act) {
#pragma unroll
      for (int i = 0; i < n; i++)
        if (i >= j) XA[i * LD + j] = Pc[i];
    }
    __syncwarp();
    double Tt[n];
#pragma unroll
    for (int i = 0; i < n; i++) Tt[i] = T[i];
    dw_solve<n, LD>(XA, invd, Tt);                             // T~[:, j]
    __syncwarp();
    // mu_s[j] = mu_f[j] + T~[:, j] . (mu_s+ - mu_p): the difference is formed per element and all-gathered
    if (DO_MEAN) {
      const double dj = musj - mupj;
      double s = mufj;
#pragma unroll
      for (int k = 0; k < n; k++) s += Tt[k] * dw_bcast(dj, k);
      musj = s;
    }
    double mtj = 0.0;
    if (DO_SIM) {
      const double dj = xsj - mupj;
      double s = mufj;
#pragma unroll
      for (int k = 0; k < n; k++) s += Tt[k] * dw_bcast(dj, k);
      mtj = s;
    }
    // T~ into XB (XB[k][i] = T~[k][i]) for the products with T~^T
    dw_put<n, LD>(XB, Tt, j, act);
    __syncwarp();
    if (DO_SIM) {
      // V~[:, j] = Sigma_f[:, j] - T~^T T[:, j];  (T~^T T)[i][j] = sum_k T~[k][i] T[k][j]
      double V[n], vinv[n];
#pragma unroll
      for (int i = 0; i < n; i++) {
        double s = Sf[i];
#pragma unroll
        for (int k = 0; k < n; k++) s -= XB[k * LD + i] * T[k];
        V[i] = act ? s : ((i == j) ? 1.0 : 0.0);
      }
      int f2 = dw_chol<n>(V, vinv, j);
      if (f2 && !fail) fail = f2;
      const double zk = act ? zb[(long)(t - 1) * n + j] : 0.0;
      double xj = 0.0;
#pragma unroll
      for (int i = 0; i < n; i++) {
        const double s = dw_sum((act && i >= j) ? V[i] * zk : 0.0);
        xj = (i == j) ? s : xj;
      }
      xsj = xj + mtj;
    }
    if (DO_VAR) {
      // G[:, j] = D T~[:, j];  Sigma_s[:, j] = Sigma_f[:, j] + T~^T G[:, j]
      double G[n];
#pragma unroll
      for (int i = 0; i < n; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < n; k++) s += XC[i * LD + k] * Tt[k];
        G[i] = s;
      }
      // Sigma_f is read again here (an L1 / L2 hit) rather than held in registers through the step
      const volatile double *hv = h;
#pragma unroll
      for (int i = 0; i < n; i++) {
        double s = (DO_SIM || !act) ? Sf[i] : hv[i * 32];
#pragma unroll
        for (int k = 0; k < n; k++) s += XB[k * LD + i] * G[k];
        Ss[i] = act ? s : 0.0;
      }
    }
    __syncwarp();
    emit(t, musj, xsj, Ss);
  }
  {  // time index 0: the initial state is known, Sigma_0 = 0
    double z0[n];
#pragma unroll
    for (int i = 0; i < n; i++) z0[i] = 0.0;
    const double v = act ? a.x0[bs * n + j] : 0.0;
    emit(0, v, v, z0);
  }
  if (LOGLIK) {
    const double lp = dw_sum(act ? lpj : 0.0);
    const double cst = -0.91893853320467274178 - log(a.gamma);
    if (valid && j == 0) a.loglik[b] = lp + cst * (double)(a.n_obs_times * a.n_obs_comp);
  }
  if (valid && j == 0 && a.status && fail && a.status[b] == 0) a.status[b] = fail;
}

template <class Fam, bool PP>
inline size_t dw_smem_bytes() {
  const size_t consts = PP ? (size_t)kDwWarps * 2 * Fam::CONST : (size_t)Fam::CONST;
  return (consts + (size_t)kDwWarps * 2 * Fam::SCR) * sizeof(double);
}

template <class Fam, class Kern>
inline cudaError_t dw_launch(Kern kern, size_t smem, const HostPrior &h, const BatchedPrior &bp, const SolveArgs &a,
                             cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  typename Fam::PriorT P;
  fill_prior(P, h);
  const unsigned grid = (unsigned)((a.B + kDwWarps * 2 - 1) / (kDwWarps * 2));
  kern<<<grid, kDwWarps * 32, smem, s>>>(P, bp, a);
  return cudaGetLastError();
}

template <class Fam, int IG>
cudaError_t launch_dw_forward(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  return dw_launch<Fam>(kalman_dw_forward<Fam, IG, false>, dw_smem_bytes<Fam, false>(), h, BatchedPrior{}, a, s);
}
template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
cudaError_t launch_dw_backward(const HostPrior &h, const SolveArgs &a, cudaStream_t s) {
  return dw_launch<Fam>(kalman_dw_backward<Fam, DO_MEAN, DO_VAR, DO_SIM, LOGLIK, false>, dw_smem_bytes<Fam, false>(), h,
                        BatchedPrior{}, a, s);
}
// per-system priors (probde interrogation)
template <class Fam>
cudaError_t launch_dw_forward_pp(const HostPrior &h, const BatchedPrior &bp, const SolveArgs &a, cudaStream_t s) {
  return dw_launch<Fam>(kalman_dw_forward<Fam, IG_PROBDE, true>, dw_smem_bytes<Fam, true>(), h, bp, a, s);
}
template <class Fam, bool DO_MEAN, bool DO_VAR, bool DO_SIM, bool LOGLIK>
cudaError_t launch_dw_backward_pp(const HostPrior &h, const BatchedPrior &bp, const SolveArgs &a, cudaStream_t s) {
  return dw_launch<Fam>(kalman_dw_backward<Fam, DO_MEAN, DO_VAR, DO_SIM, LOGLIK, true>, dw_smem_bytes<Fam, true>(), h, bp,
                        a, s);
}

}  // namespace rodeo
