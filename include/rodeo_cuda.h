/*
 * rodeo_cuda.h -- C ABI of the B200-native batched KalmanODE backend.
 *
 * This is the drop-in boundary for the ONE hot path of mlysy/rodeo-legacy: a batch of
 * independent KalmanODE solves (forward time-varying Kalman filter + backward
 * smoother).  Plain pointers and sizes only; no torch / C++ types.  A binding in the
 * reference (Cython `cdef extern`, ctypes, cffi) needs nothing but this header and
 * librodeo_cuda.so -- see INTEGRATION.md.
 *
 * What each entry point replaces in the reference (paths relative to the reference):
 *
 *   rodeo_cuda_problem_create   KalmanODE.__cinit__            rodeo/cython/KalmanODE.pyx:170-218
 *                               KalmanTVODE::KalmanTVODE       rodeo/eigen/KalmanTVODE.h:104-139
 *                               (cdef extern at rodeo/eigen/KalmanTVODE.pxd:1-14)
 *   rodeo_cuda_problem_set      the wgt_meas/mu_state/wgt_state/var_state property setters
 *                                                              rodeo/cython/KalmanODE.pyx:220-270
 *                               and the per-call `W` override  rodeo/cython/KalmanODE.pyx:441-442
 *   rodeo_cuda_solve            _solve_filter + solve_mv / solve_sim / solve, batched
 *                                                              rodeo/cython/KalmanODE.pyx:280-463
 *                               predict/forecast_probde/update/smooth_update
 *                                                              rodeo/eigen/KalmanTVODE.h:149-183, 273-296, 333-341
 *                               with the arithmetic of         rodeo/eigen/KalmanTV.h:280-294, 324-357,
 *                                                              447-467, 503-533, 619-628
 *   rodeo_cuda_solve_host       the same, with HOST buffers (what a NumPy caller of the
 *                               reference passes): H2D / D2H copies are done inside.
 *   rodeo_cuda_loglik           kalman_nlpost's solve -> thinning -> loglike chain
 *                                                              examples/inference.py:54-56, 66-94
 *
 * Conventions
 *   - float64 only (as the reference, README.md:15-16).
 *   - Prior / measurement matrices are COLUMN-MAJOR (Fortran order), exactly what the
 *     reference's constructor receives: wgt_meas (n_meas x n_state), wgt_state and
 *     var_state (n_state x n_state), mu_state (n_state).
 *   - Batched I/O is batch-major C order:  x0 (B, n_state), theta (B, n_theta),
 *     x_out / mu_out (B, n_eval+1, n_state), var_out (B, n_eval+1, n_state, n_state):
 *     per system these are the reference's return values (the `.T` views at
 *     KalmanODE.pyx:394, 427, 463).  var_out[:, 0] is written as zeros (the reference
 *     leaves it uninitialised; the initial state is known exactly).
 *   - z (standard-normal draws): column-major (n_state x n_eval) per system as the
 *     reference's z_state; column k is consumed at time index k+1
 *     (KalmanODE.pyx:410-425).  z_batch_stride = 0 shares one z across the batch,
 *     n_state*n_eval gives every system its own.
 *   - ode_fun is a registered device functor selected by id (there is no Python
 *     callback on the path): RHS formulas follow tests/eigen/ode_functions.pyx:8-60
 *     and examples/{chkrebtii,fitz,lorenz,mseir}.py.
 *   - All functions return 0 on success, a negative RODEO_E_* code otherwise;
 *     rodeo_cuda_last_error() gives the message (thread-local).  Per-system Cholesky
 *     failures (non-positive pivot; silently ignored by the reference) are reported
 *     in `status` (0 = ok, k+1 = first failed pivot index) and do not fail the call.
 *   - A handle is not thread-safe; one solve in flight per handle.  Different handles may be
 *     created and used from different threads.  rodeo_cuda_problem_destroy leaves the calling
 *     thread's current CUDA device as it found it.
 */
#ifndef RODEO_CUDA_H
#define RODEO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RODEO_CUDA_ABI_VERSION 1

/* registered right-hand sides */
enum {
  RODEO_ODE_CHKREBTII = 0, /* x'' = sin(2t) - x          n_meas 1, n_theta 0 */
  RODEO_ODE_FITZ = 1,      /* FitzHugh-Nagumo            n_meas 2, n_theta 3 */
  RODEO_ODE_LORENZ63 = 2,  /* Lorenz 63                  n_meas 3, n_theta 3 */
  RODEO_ODE_MSEIR = 3      /* MSEIR epidemic             n_meas 5, n_theta 6 */
};

/* what the backward pass returns (KalmanODE.solve_mv / solve_sim / solve) */
enum { RODEO_MODE_MV = 1, RODEO_MODE_SIM = 2, RODEO_MODE_BOTH = 3 };

/* interrogation variants (rodeo/cython/KalmanODE.pyx:13-102); probde is what ships */
enum { RODEO_INTERROGATE_PROBDE = 0, RODEO_INTERROGATE_SCHOBERT = 1, RODEO_INTERROGATE_CHKREBTII = 2 };

/* fields selectable in rodeo_cuda_problem_set */
enum { RODEO_FIELD_WGT_MEAS = 0, RODEO_FIELD_MU_STATE = 1, RODEO_FIELD_WGT_STATE = 2, RODEO_FIELD_VAR_STATE = 3 };

/* problem flags */
enum {
  RODEO_FLAG_FORCE_DENSE = 1,  /* never use the block-diagonal kernel family, even when Q, R, W allow it */
  RODEO_FLAG_NO_STRUCTURE = 2  /* block-diagonal family without the upper-triangular-Q / unit-w specialisation */
};

/* error codes */
enum {
  RODEO_OK = 0,
  RODEO_E_INVALID = -1,     /* bad argument (NULL, non-positive size, unknown id) */
  RODEO_E_UNSUPPORTED = -2, /* (ode, n_state, n_meas, interrogate) has no compiled kernel */
  RODEO_E_CUDA = -3,        /* CUDA runtime error (message has the CUDA string) */
  RODEO_E_NOMEM = -4        /* workspace allocation failed */
};

typedef struct rodeo_problem rodeo_problem; /* opaque */

typedef struct {
  int32_t n_state;        /* n */
  int32_t n_meas;         /* m */
  int32_t n_eval;         /* N: number of steps; outputs have N+1 time points */
  int32_t ode;            /* RODEO_ODE_* */
  int32_t interrogate;    /* RODEO_INTERROGATE_* */
  int32_t flags;          /* RODEO_FLAG_* or 0 */
  double tmin, tmax;
  const double *wgt_meas;  /* host, column-major (m x n) */
  const double *wgt_state; /* host, column-major (n x n) */
  const double *mu_state;  /* host, (n) */
  const double *var_state; /* host, column-major (n x n), symmetric */
} rodeo_problem_desc;

/* ---- lifetime ------------------------------------------------------------------- */
int rodeo_cuda_abi_version(void);
const char *rodeo_cuda_last_error(void);

/* n_meas / n_theta of a registered RHS, or RODEO_E_INVALID; ids run 0 .. rodeo_cuda_ode_count() - 1
 * (all three come from the compiled device functors, csrc/ode_functors.cuh) */
int rodeo_cuda_ode_count(void);
int rodeo_cuda_ode_n_meas(int ode);
int rodeo_cuda_ode_n_theta(int ode);
/* Bit 0: a dense kernel set is compiled for this combination (any Q, R, W).  Bit 1: a
 * block-diagonal set is compiled (used automatically when wgt_state / var_state are block
 * diagonal with n_meas equal blocks and row a of wgt_meas lies inside block a -- what
 * ibm_init / car_init + indep_init + zero_pad always produce).  0: unsupported. */
int rodeo_cuda_supported(int ode, int n_state, int n_meas, int interrogate);
/* The same question for a list of block sizes (the reference's n_deriv_prior, one entry per ODE variable:
 * rodeo/utils/utils.py:97-105, rodeo/ibm/ibm_init.py:54-58 accept any list).  Equal sizes answer like
 * rodeo_cuda_supported; UNEQUAL sizes (e.g. [3, 4]) return 1 when a dense kernel set was compiled for exactly that
 * list (csrc/kernels_mixed.cu), else 0.  A handle locates the variables from wgt_meas: when every row is a unit
 * vector e_c (zero_pad), variable a's block starts at c - (order of the ODE); otherwise the reference functors'
 * equal split n_state / n_meas is used (tests/eigen/ode_functions.pyx:33-42). */
int rodeo_cuda_supported_blocks(int ode, int n_blocks, const int32_t *block_sizes, int interrogate);
/* Enumerate every compiled kernel set, index 0 .. rodeo_cuda_kernel_set_count() - 1: right-hand side id, n_state,
 * family (blockdiag / structured) and the block sizes it was compiled for (n_meas entries written to block_sizes;
 * any out pointer may be NULL).  *name points at a static string such as "fitz/n7[3,4]". */
int rodeo_cuda_kernel_set_count(void);
int rodeo_cuda_kernel_set_info(int index, int32_t *ode, int32_t *n_state, int32_t *blockdiag, int32_t *structured,
                               int32_t *block_sizes, const char **name);

/* Creates a problem on the CURRENT CUDA device.  The descriptor's host arrays are copied. */
int rodeo_cuda_problem_create(const rodeo_problem_desc *desc, rodeo_problem **out);
int rodeo_cuda_problem_destroy(rodeo_problem *p);
/* Overwrite one of the prior / measurement arrays (host pointer, column-major). */
int rodeo_cuda_problem_set(rodeo_problem *p, int field, const double *host_values);

/* Which kernel family the handle currently uses (it is re-selected whenever a field is set).
 * *registers: 1 = thread-per-system register kernels, 2 = dense wide-state kernels (half a warp per
 * system, n_state 7..16), 0 = rolled local-memory fallback. */
int rodeo_cuda_problem_info(const rodeo_problem *p, int32_t *blockdiag, int32_t *registers,
                            int32_t *hist_doubles_per_step);
/* 1 when the block-diagonal kernels in use are the ones specialised for structured priors: every Q block
 * upper triangular and every wgt_meas row the unit vector of the ODE's order inside its block (what ibm_init
 * + zero_pad produce).  They skip the exactly-zero terms: same results bit for bit, ~40 % fewer FP64
 * instructions in the filter step. */
int rodeo_cuda_problem_structured(const rodeo_problem *p);

/* ---- workspace (filter history streamed through HBM between the two passes) ------ */
/* Bytes of history one system needs for this problem. */
size_t rodeo_cuda_history_bytes_per_system(const rodeo_problem *p);
/* Upper bound for the handle-owned history workspace (default: 48 GiB).  A batch whose
 * history exceeds it is processed in chunks, transparently. */
int rodeo_cuda_set_workspace_limit(rodeo_problem *p, size_t bytes);
/* Checkpoint interval of the filter history: 1 stores the filtered moments of every time
 * index; k = 2 or 4 stores every k-th and the backward pass re-runs the filter in between
 * (bit-identical results, k-fold less history traffic and workspace); 0 (default) = auto: 2
 * where a kernel exists for it.  Honoured by the thread-per-system register kernels with the
 * probde interrogation; otherwise 1 is used.
 * rodeo_cuda_get_checkpoint returns the interval actually in effect. */
int rodeo_cuda_set_checkpoint(rodeo_problem *p, int interval);
int rodeo_cuda_get_checkpoint(const rodeo_problem *p);
/* Shared-covariance mode.  One handle has one prior (Q, c, R, W), so the covariance path
 * (Sigma_pred / filt / smooth, Kalman and smoother gains, sampling Cholesky factors) is the same
 * for every system of a batch: it does not depend on theta, x0 or the ODE (rodeo/eigen/KalmanTV.h
 * never reads a mean when producing a variance).  With this mode on, those are computed ONCE per
 * prior into device tables and the batch kernels run the mean recursions only; `var_out` of
 * rodeo_cuda_solve / rodeo_cuda_solve_host then receives ONE (n_eval+1, n_state, n_state) array
 * that holds for all systems, instead of one per system.  Means, draws and log-likelihoods agree
 * with the general mode to FP64 rounding.  Off by default: it is a different amount of work per
 * system than the general mode and is benchmarked separately.  RODEO_E_UNSUPPORTED for kernel
 * families without it (rolled dense n > 6) and for the chkrebtii interrogation. */
int rodeo_cuda_set_shared_cov(rodeo_problem *p, int enabled);
int rodeo_cuda_get_shared_cov(const rodeo_problem *p);
/* Layout of `var_out` (opt-in; the default RODEO_VAR_FULL is the reference's (n_state, n_state) matrix per
 * system and time index).  For block-diagonal priors 2/3 (FitzHugh-Nagumo) to 4/5 (MSEIR) of that matrix are
 * structural zeros and it is 82 % of the bytes a solve_mv writes and returns:
 *   RODEO_VAR_BLOCKS  var_out (B, n_eval+1, n_meas, p, p): the diagonal blocks only, block a = Sigma[a*p:(a+1)*p, a*p:(a+1)*p]
 *   RODEO_VAR_DIAG    var_out (B, n_eval+1, n_state): the marginal variances only
 * Same numbers as the corresponding entries of the full output, bit for bit.  Available for block-diagonal priors
 * with at least two blocks in general mode; RODEO_E_UNSUPPORTED otherwise (dense priors, single-block priors).  It applies to
 * rodeo_cuda_solve / rodeo_cuda_solve_host (not to the batched-prior entry point, nor to shared-covariance mode). */
enum { RODEO_VAR_FULL = 0, RODEO_VAR_BLOCKS = 1, RODEO_VAR_DIAG = 2 };
int rodeo_cuda_set_var_layout(rodeo_problem *p, int layout);
int rodeo_cuda_get_var_layout(const rodeo_problem *p);
/* doubles of var_out per (system, time index) in the current layout: n*n, n_meas*p*p or n */
int rodeo_cuda_var_doubles_per_step(const rodeo_problem *p);
/* Pre-allocate the workspace for batches up to max_batch (otherwise grown on demand). */
int rodeo_cuda_reserve(rodeo_problem *p, int64_t max_batch);

/* ---- the hot path ----------------------------------------------------------------- */
/* Device-pointer entry point.  All pointers are device memory on the problem's device.
 * theta may be NULL when the RHS has no parameters; z may be NULL for RODEO_MODE_MV with
 * probde/schobert interrogation; outputs not produced by `mode` may be NULL; status may
 * be NULL.  Output pointers must be 16-byte aligned (they are written with vector / TMA
 * stores; RODEO_E_INVALID otherwise).  Work is enqueued on `cuda_stream` (a cudaStream_t,
 * NULL = default stream) and the call returns without synchronising. */
int rodeo_cuda_solve(rodeo_problem *p, int mode, int64_t batch, const double *x0,
                     const double *theta, const double *z, int64_t z_batch_stride,
                     double *x_out, double *mu_out, double *var_out, int32_t *status,
                     void *cuda_stream);

/* Per-system priors (e.g. a sweep over the prior scale sigma).  As rodeo_cuda_solve,
 * but every system may bring its own wgt_meas (m x n), mu_state (n), wgt_state (n x n),
 * var_state (n x n): device arrays, batch-major, each system column-major like the constructor's
 * arrays.  A NULL array means "the handle's shared value for every system".  When the handle's prior and every
 * system's arrays are block diagonal (two equal blocks: the FitzHugh-Nagumo family) the thread-per-block kernels
 * run with a per-lane prior (checkpoint 2, granule stores); otherwise the dense kernel family with the full
 * history (each thread keeps its prior in registers / local memory instead of the constant bank).  probde
 * interrogation only.  Synchronises `cuda_stream` once (structure check). */
int rodeo_cuda_solve_batched_prior(rodeo_problem *p, int mode, int64_t batch, const double *x0,
                                   const double *theta, const double *z, int64_t z_batch_stride,
                                   const double *wgt_meas_b, const double *mu_state_b,
                                   const double *wgt_state_b, const double *var_state_b, double *x_out,
                                   double *mu_out, double *var_out, int32_t *status, void *cuda_stream);

/* 1 when the most recent rodeo_cuda_solve_batched_prior call ran on the block-diagonal thread-per-block kernels:
 * the handle's prior and EVERY system's own arrays are block diagonal with n_meas equal blocks and row a of W
 * inside block a (checked on the device per call; e.g. a sweep over the IBM scale sigma).  0: dense family. */
int rodeo_cuda_last_batched_prior_blockdiag(const rodeo_problem *p);

/* rodeo_cuda_loglik with per-system priors (a sigma-tuning likelihood sweep without materialising trajectories):
 * the batched arrays as in rodeo_cuda_solve_batched_prior, the observation arguments as in rodeo_cuda_loglik.
 * Dense kernel family, probde interrogation. */
int rodeo_cuda_loglik_batched_prior(rodeo_problem *p, int use_draw, int64_t batch, const double *x0,
                                    const double *theta, const double *z, int64_t z_batch_stride,
                                    const double *wgt_meas_b, const double *mu_state_b,
                                    const double *wgt_state_b, const double *var_state_b,
                                    const int32_t *thin_index, int32_t n_obs_times, const int32_t *state_index,
                                    int32_t n_obs_comp, const double *Y, double gamma, double *loglik_out,
                                    int32_t *status, void *cuda_stream);

/* Host-pointer entry point (the call a NumPy user of the reference makes): copies the
 * inputs to the device, solves in chunks with copy/compute overlap, copies the results
 * back and synchronises.  Host buffers may be pageable or pinned (pinned is faster). */
int rodeo_cuda_solve_host(rodeo_problem *p, int mode, int64_t batch, const double *x0,
                          const double *theta, const double *z, int64_t z_batch_stride,
                          double *x_out, double *mu_out, double *var_out, int32_t *status);

/* Thinned Gaussian log-likelihood of the posterior mean (use_draw = 0) or of a posterior
 * draw (use_draw = 1, needs z) against observations Y (n_obs_times x n_obs_comp, row
 * major):  sum_d sum_k  log N(Y[d,k]; X[thin_index[d], state_index[k]], sd = gamma).
 * One double per system in loglik_out (device).  No per-step output is materialised. */
int rodeo_cuda_loglik(rodeo_problem *p, int use_draw, int64_t batch, const double *x0,
                      const double *theta, const double *z, int64_t z_batch_stride,
                      const int32_t *thin_index, int32_t n_obs_times,
                      const int32_t *state_index, int32_t n_obs_comp, const double *Y,
                      double gamma, double *loglik_out, int32_t *status, void *cuda_stream);

/* ---- introspection for benchmarks -------------------------------------------------- */
/* kernels launched by this handle since creation (or since the last reset) */
int64_t rodeo_cuda_launch_count(const rodeo_problem *p);
int rodeo_cuda_reset_launch_count(rodeo_problem *p);
/* Device time (ms) of the forward and backward kernels of the most recent
 * rodeo_cuda_solve call, measured with CUDA events on its stream; synchronises. */
int rodeo_cuda_last_kernel_ms(rodeo_problem *p, float *forward_ms, float *backward_ms);
/* The same summed over EVERY solve since rodeo_cuda_set_timing(p, 1) (at most 4096 chunks are kept),
 * with the number of solves: bench.py divides by it for the per-step kernel times; synchronises. */
int rodeo_cuda_kernel_ms_total(rodeo_problem *p, double *forward_ms, double *backward_ms, int32_t *n_solves);
/* test hook: copy the first `bytes` of the handle's history workspace to a device buffer.  Layout
 * hist[tile][record][element][32 systems] (csrc/kalman_common.cuh); for two-block block-diagonal priors
 * solved by the thread-per-block kernels, system b sits at position perm16(b) (same file). */
int rodeo_cuda_debug_copy_history(rodeo_problem *p, void *dst_device, size_t bytes, void *cuda_stream);
/* test hook: the thread-per-block kernels address a 16-system output row with 32-bit byte offsets; problems whose
 * rows reach `bytes` (default and maximum 2^31: ~466,000 time steps at n_state 6) run on the previous-generation
 * kernels instead.  Lowering it exercises that hand-over on ordinary sizes. */
int rodeo_cuda_debug_set_row_limit(rodeo_problem *p, double bytes);
/* enable (1) / disable (0) the event recording used by rodeo_cuda_last_kernel_ms / _kernel_ms_total
 * (either value restarts the accumulation) */
int rodeo_cuda_set_timing(rodeo_problem *p, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* RODEO_CUDA_H */
