// registry.cuh -- table of compiled kernel sets and their launchers.
#pragma once
#include <cuda_runtime.h>

#include "kalman_tps.cuh"

namespace rodeo {

// Everything the host side needs to fill a Prior / PriorBD without knowing n, m statically.
struct HostPrior {
  int n, m, n_eval;
  double tmin, tmax;
  const double *W, *Q, *c, *R;  // column-major host arrays
};

enum BackwardKind { BK_MV = 0, BK_SIM = 1, BK_BOTH = 2, BK_LL_MEAN = 3, BK_LL_DRAW = 4, BK_COUNT = 5 };

constexpr int kNumCk = 3;                       // checkpoint intervals 1, 2, 4
inline int ck_slot(int ck) { return ck == 1 ? 0 : ck == 2 ? 1 : ck == 4 ? 2 : -1; }

typedef cudaError_t (*launch_fn)(const HostPrior &, const SolveArgs &, cudaStream_t);
// Per-system priors: device arrays, batch-major, each system laid out like the constructor's
// arrays (column-major); nullptr = use the handle's shared value.
struct BatchedPrior {
  const double *W, *Q, *c, *R;
};
typedef cudaError_t (*launch_pp_fn)(const HostPrior &, const BatchedPrior &, const SolveArgs &, cudaStream_t);

struct KernelSet {
  int ode, n, m;
  bool blockdiag;                 // block-diagonal family (needs structured Q, R, W) or dense
  bool registers;                 // fully unrolled register kernels (else rolled / local memory)
  bool structured;                // block-diagonal family specialised for upper-triangular Q blocks and w = e_order
  bool mixed;                     // dense set compiled for UNEQUAL block sizes (OdeAt, ode_functors.cuh): the
  int off[8];                     //   right-hand side reads variable k at state index off[k] (equal split: k * n / m)
  const char *name;
  launch_fn forward[3];           // by interrogation variant
  launch_fn forward_perm[3];      // forward pass that writes the permuted history order (nullptr: forward[] honours perm16)
  launch_fn backward[kNumCk][BK_COUNT];  // by checkpoint slot and BackwardKind (nullptr: not compiled)
  // the previous generation of a kernel where one is kept for A/B measurements (RODEO_CUDA_BACKWARD=staged)
  launch_fn backward_alt[kNumCk][BK_COUNT];
  bool backward_perm[kNumCk][BK_COUNT];
  // compact variance layouts (rodeo_cuda_set_var_layout): [layout - 1][0 = solve_mv, 1 = solve]; nullptr when not
  // compiled for this set (two blocks: thread-per-block kernels; >= 3 blocks: warp-per-block kernels)
  launch_fn backward_vl[2][2];
  int vl_ck;                      // ... the checkpoint interval those launchers use, and whether they want the
  bool vl_perm;                   //     permuted history order  // backward[][] wants the history in the permuted order (SolveArgs::perm16)
  int hist_elems;                 // doubles of history per system per step
  // shared-covariance mode (kalman_sc.cuh); all nullptr when not compiled for this set
  cudaError_t (*sc_tables)(const HostPrior &, int zero_H, double *tab, int *fail_out, cudaStream_t);
  long (*sc_table_doubles)(int N);
  long (*sc_table_var_offset)(int N);
  launch_fn sc_forward;
  launch_fn sc_backward[BK_COUNT];
  // per-system priors (dense family, probde interrogation); nullptr when not compiled
  launch_pp_fn pp_forward;
  launch_pp_fn pp_backward[BK_COUNT];
  // ... for two-block block-diagonal sets when every system's prior is block diagonal too: the thread-per-block
  // kernels with a per-lane prior (checkpoint 2, permuted history order); nullptr when not compiled
  launch_pp_fn pp_bd_forward;
  launch_pp_fn pp_bd_backward[BK_COUNT];
  // dense wide-state kernels (kalman_dw.cuh: half a warp per system); nullptr / 0 when not compiled.
  // Used for every entry point under the probde and schobert interrogations; they keep
  // their own history layout (wide_hist_elems doubles per system per step, full history).
  launch_fn wide_forward[3];           // by interrogation variant
  launch_fn wide_backward[BK_COUNT];
  launch_pp_fn wide_pp_forward;
  launch_pp_fn wide_pp_backward[BK_COUNT];
  int wide_hist_elems;
};

int ode_count();                 // registered right-hand sides: ids 0 .. ode_count() - 1
int ode_n_meas(int ode);         // from the functors (ode_functors.cuh); -1 for an unknown id
int ode_n_theta(int ode);
int ode_order(int ode);          // order of the ODE: the derivative zero_pad's wgt_meas picks in every block
// off = nullptr: the equal-split sets; else the dense set compiled for exactly these variable offsets
const KernelSet *find_kernel_set(int ode, int n, bool blockdiag, bool structured = false, const int *off = nullptr);
int kernel_set_count();
const KernelSet *kernel_set_at(int i);

}  // namespace rodeo
