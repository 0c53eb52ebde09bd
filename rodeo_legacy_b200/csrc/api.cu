// api.cu -- the C ABI declared in include/rodeo_cuda.h.
//
// Host-side runtime around the kernels: problem handles, the HBM history workspace,
// batch chunking, the pinned/pageable host path with copy/compute overlap, launch
// accounting and event timing.  No torch, no Python: plain CUDA runtime.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rodeo_cuda.h"
#include "registry.cuh"

using namespace rodeo;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
static int cuda_fail(cudaError_t e, const char *what) {
  return fail(RODEO_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                      \
  do {                                                \
    cudaError_t e_ = (call);                          \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

struct rodeo_problem {
  int n, m, N, ode, ig, device;
  double tmin, tmax;
  std::vector<double> W, Q, c, R;  // column-major host copies
  const KernelSet *ks;
  bool mixed = false;              // unequal block sizes: wgt_meas places the variables at off[] (a "mixed" dense set)
  int off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int flags = 0;
  int *bd_flag_dev = nullptr;      // device flag of the batched-prior structure check
  bool last_pp_bd = false;         // the last batched-prior solve ran on the block-diagonal kernels
  int ck = 0;                      // requested checkpoint interval of the filter history (0 = auto)
  int var_layout = 0;              // RODEO_VAR_FULL / _BLOCKS / _DIAG
  double row_limit = 2147483648.0; // bytes of a 16-system output row the thread-per-block kernels address (32-bit)
  // shared-covariance mode (kalman_sc.inl): tables cached per (prior, N), rebuilt after problem_set
  bool shared = false;
  double *tab = nullptr;
  size_t tab_bytes = 0;
  bool tab_valid = false;
  int tab_fail = 0;
  int *tab_fail_dev = nullptr;
  void *ws = nullptr;
  size_t ws_bytes = 0;
  size_t ws_limit = (size_t)48 << 30;
  int64_t launches = 0;
  bool timing = false;
  std::vector<cudaEvent_t> ev;     // 3 per chunk, every chunk launched since timing was switched on
  int ev_first = 0;                // first chunk of the LAST solve in ev (chunk index)
  int ev_chunks = 0;               // chunks recorded since timing was switched on
  int ev_solves = 0;               // solves recorded since timing was switched on
  // host-path device buffers (grow-only)
  void *hbuf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t hbuf_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t e_done[2] = {nullptr, nullptr}, e_copied[2] = {nullptr, nullptr};
};


// Q, R block diagonal with m blocks of size n/m and row a of W supported inside block a?
static bool is_block_diagonal(const rodeo_problem *p) {
  const int n = p->n, m = p->m;
  if (n % m) return false;
  const int bs = n / m;
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++)
      if (i / bs != j / bs && (p->Q[i + (size_t)j * n] != 0.0 || p->R[i + (size_t)j * n] != 0.0)) return false;
  for (int j = 0; j < n; j++)
    for (int a = 0; a < m; a++)
      if (j / bs != a && p->W[a + (size_t)j * m] != 0.0) return false;
  return true;
}

// A block-diagonal prior whose Q blocks are all upper triangular and whose W rows are the unit vector
// e_order inside their block (ibm_init + zero_pad): the structured kernel sets skip the exact zeros.
static bool is_structured(const rodeo_problem *p) {
  const int n = p->n, m = p->m, bs = n / m, ord = ode_order(p->ode);
  if (ord < 0 || ord >= bs) return false;
  for (int b = 0; b < m; b++) {
    for (int j = 0; j < bs; j++) {
      for (int i = j + 1; i < bs; i++)
        if (p->Q[(b * bs + i) + (size_t)(b * bs + j) * n] != 0.0) return false;      // strictly lower part
      if (p->W[b + (size_t)(b * bs + j) * m] != (j == ord ? 1.0 : 0.0)) return false;
    }
  }
  return true;
}

// Where wgt_meas puts the ODE variables: when every row a is a unit vector e_c (what zero_pad produces,
// rodeo/utils/utils.py:84-105), variable a's block starts at c - order.  False when W is not of that form (a general
// measurement matrix: the right-hand side then reads the reference's equal split, tests/eigen/ode_functions.pyx:33-42).
static bool offsets_from_wgt_meas(const rodeo_problem *p, int *off) {
  const int n = p->n, m = p->m, ord = ode_order(p->ode);
  if (m > 8) return false;
  for (int a = 0; a < m; a++) {
    int col = -1;
    for (int j = 0; j < n; j++) {
      const double w = p->W[a + (size_t)j * m];
      if (w == 0.0) continue;
      if (w != 1.0 || col >= 0) return false;
      col = j;
    }
    if (col < ord) return false;
    off[a] = col - ord;
    if (a == 0 ? off[0] != 0 : off[a] <= off[a - 1]) return false;
  }
  return true;
}

// the dense kernel set that serves this problem's shape (per-system priors, forced-dense runs)
static const KernelSet *dense_set(const rodeo_problem *p) {
  return p->mixed ? find_kernel_set(p->ode, p->n, false, false, p->off) : find_kernel_set(p->ode, p->n, false);
}

// Pick the kernel family for the current (Q, R, W): block-diagonal when the structure is
// there (and not disabled by RODEO_FLAG_FORCE_DENSE), dense otherwise.
static int select_kernel_set(rodeo_problem *p) {
  int off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool equal = true;
  if (offsets_from_wgt_meas(p, off)) {
    for (int a = 0; a < p->m; a++) equal = equal && p->n % p->m == 0 && off[a] == a * (p->n / p->m);
  } else if (p->n % p->m) {
    return fail(RODEO_E_UNSUPPORTED, "no compiled kernel: n_state is not a multiple of n_meas and wgt_meas does not "
                                     "locate the variables (its rows must be unit vectors, as zero_pad builds them)");
  }
  p->mixed = !equal;
  for (int a = 0; a < 8; a++) p->off[a] = equal ? 0 : off[a];
  if (p->mixed) {   // unequal block sizes: a dense set compiled for exactly these variable offsets
    p->ks = dense_set(p);
    if (p->ks) return RODEO_OK;
    char buf[320], sz[96] = "";
    for (int a = 0, len = 0; a < p->m; a++)
      len += snprintf(sz + len, sizeof sz - len, "%s%d", a ? ", " : "", (a + 1 < p->m ? off[a + 1] : p->n) - off[a]);
    snprintf(buf, sizeof buf, "no compiled kernel for ode=%d with block sizes [%s] (n_state=%d); compiled shapes: "
             "rodeo_cuda_kernel_set_info, adding one: INTEGRATION.md section 4.1", p->ode, sz, p->n);
    return fail(RODEO_E_UNSUPPORTED, buf);
  }
  const KernelSet *dense = find_kernel_set(p->ode, p->n, false), *bd = find_kernel_set(p->ode, p->n, true);
  const bool structured = is_block_diagonal(p);
  const KernelSet *ks = nullptr;
  if (bd && structured && !(p->flags & RODEO_FLAG_FORCE_DENSE)) {
    const KernelSet *bds = find_kernel_set(p->ode, p->n, true, true);
    ks = (bds && !(p->flags & RODEO_FLAG_NO_STRUCTURE) && is_structured(p)) ? bds : bd;
  } else if (dense) ks = dense;
  if (!ks) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "no compiled kernel for ode=%d n_state=%d n_meas=%d%s (see rodeo_cuda_supported)", p->ode, p->n,
             p->m, bd ? " with a prior that is not block diagonal" : "");
    return fail(RODEO_E_UNSUPPORTED, buf);
  }
  p->ks = ks;
  return RODEO_OK;
}

extern "C" {

int rodeo_cuda_abi_version(void) { return RODEO_CUDA_ABI_VERSION; }
const char *rodeo_cuda_last_error(void) { return g_err.c_str(); }

int rodeo_cuda_ode_n_meas(int ode) { return ode_n_meas(ode) >= 0 ? ode_n_meas(ode) : RODEO_E_INVALID; }
int rodeo_cuda_ode_n_theta(int ode) { return ode_n_theta(ode) >= 0 ? ode_n_theta(ode) : RODEO_E_INVALID; }
int rodeo_cuda_ode_count(void) { return ode_count(); }

// 1: a dense kernel set exists (any Q, R, W); 2: only a block-diagonal set exists (the prior
// must be block diagonal with n_meas blocks and W row a inside block a); 3: both; 0: none.
int rodeo_cuda_supported(int ode, int n_state, int n_meas, int interrogate) {
  if (interrogate < 0 || interrogate > 2) return 0;
  const KernelSet *d = find_kernel_set(ode, n_state, false), *b = find_kernel_set(ode, n_state, true);
  return ((d && d->m == n_meas) ? 1 : 0) | ((b && b->m == n_meas) ? 2 : 0);
}

// The same question for a list of block sizes (n_deriv_prior): equal sizes answer like rodeo_cuda_supported,
// unequal sizes 1 when a dense set was compiled for exactly that list, else 0.
int rodeo_cuda_supported_blocks(int ode, int n_blocks, const int32_t *block_sizes, int interrogate) {
  if (!block_sizes || n_blocks <= 0 || n_blocks > 8 || n_blocks != ode_n_meas(ode)) return 0;
  int off[8], n = 0;
  bool equal = true;
  for (int a = 0; a < n_blocks; a++) {
    if (block_sizes[a] <= 0) return 0;
    off[a] = n;
    n += block_sizes[a];
    equal = equal && block_sizes[a] == block_sizes[0];
  }
  if (equal) return rodeo_cuda_supported(ode, n, n_blocks, interrogate);
  if (interrogate < 0 || interrogate > 2) return 0;
  return find_kernel_set(ode, n, false, false, off) ? 1 : 0;
}

// Enumerate the compiled kernel sets: index 0 .. rodeo_cuda_kernel_set_count() - 1.  block_sizes gets n_meas entries.
int rodeo_cuda_kernel_set_count(void) { return kernel_set_count(); }
int rodeo_cuda_kernel_set_info(int index, int32_t *ode, int32_t *n_state, int32_t *blockdiag, int32_t *structured,
                               int32_t *block_sizes, const char **name) {
  const KernelSet *k = kernel_set_at(index);
  if (!k) return fail(RODEO_E_INVALID, "kernel set index out of range");
  if (ode) *ode = k->ode;
  if (n_state) *n_state = k->n;
  if (blockdiag) *blockdiag = k->blockdiag;
  if (structured) *structured = k->structured;
  if (block_sizes)
    for (int a = 0; a < k->m; a++) block_sizes[a] = (a + 1 < k->m ? k->off[a + 1] : k->n) - k->off[a];
  if (name) *name = k->name;
  return RODEO_OK;
}

int rodeo_cuda_problem_create(const rodeo_problem_desc *d, rodeo_problem **out) {
  if (!d || !out) return fail(RODEO_E_INVALID, "null descriptor or output");
  *out = nullptr;
  if (d->n_state <= 0 || d->n_meas <= 0 || d->n_eval < 1)
    return fail(RODEO_E_INVALID, "n_state, n_meas must be positive and n_eval >= 1");
  if (!d->wgt_meas || !d->wgt_state || !d->mu_state || !d->var_state)
    return fail(RODEO_E_INVALID, "wgt_meas, wgt_state, mu_state, var_state must be set");
  if (ode_n_meas(d->ode) < 0) return fail(RODEO_E_INVALID, "unknown ode id");
  if (d->interrogate < 0 || d->interrogate > 2) return fail(RODEO_E_INVALID, "unknown interrogate id");
  if (d->n_meas != ode_n_meas(d->ode)) return fail(RODEO_E_INVALID, "n_meas does not match the ode");
  rodeo_problem *p = new rodeo_problem();
  p->n = d->n_state;
  p->m = d->n_meas;
  p->N = d->n_eval;
  p->ode = d->ode;
  p->ig = d->interrogate;
  p->device = 0;
  p->tmin = d->tmin;
  p->tmax = d->tmax;
  p->flags = d->flags;
  p->W.assign(d->wgt_meas, d->wgt_meas + (size_t)p->m * p->n);
  p->Q.assign(d->wgt_state, d->wgt_state + (size_t)p->n * p->n);
  p->c.assign(d->mu_state, d->mu_state + p->n);
  p->R.assign(d->var_state, d->var_state + (size_t)p->n * p->n);
  int rc = select_kernel_set(p);   // before any CUDA call: unsupported shapes fail the same way everywhere
  if (rc) {
    delete p;
    return rc;
  }
  cudaError_t ce = cudaGetDevice(&p->device);
  if (ce != cudaSuccess) {
    delete p;
    return cuda_fail(ce, "cudaGetDevice");
  }
  *out = p;
  return RODEO_OK;
}

int rodeo_cuda_problem_destroy(rodeo_problem *p) {
  if (!p) return RODEO_OK;
  int prev = -1;
  cudaGetDevice(&prev);            // a garbage-collected handle must not move the caller's current device
  cudaSetDevice(p->device);
  if (p->ws) cudaFree(p->ws);
  if (p->tab) cudaFree(p->tab);
  if (p->tab_fail_dev) cudaFree(p->tab_fail_dev);
  if (p->bd_flag_dev) cudaFree(p->bd_flag_dev);
  for (int i = 0; i < 8; i++)
    if (p->hbuf[i]) cudaFree(p->hbuf[i]);
  for (auto e : p->ev) cudaEventDestroy(e);
  for (int i = 0; i < 2; i++) {
    if (p->e_done[i]) cudaEventDestroy(p->e_done[i]);
    if (p->e_copied[i]) cudaEventDestroy(p->e_copied[i]);
  }
  if (p->s_compute) cudaStreamDestroy(p->s_compute);
  if (p->s_copy) cudaStreamDestroy(p->s_copy);
  if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
  delete p;
  return RODEO_OK;
}

int rodeo_cuda_problem_set(rodeo_problem *p, int field, const double *v) {
  if (!p || !v) return fail(RODEO_E_INVALID, "null problem or values");
  switch (field) {
    case RODEO_FIELD_WGT_MEAS: p->W.assign(v, v + (size_t)p->m * p->n); break;
    case RODEO_FIELD_MU_STATE: p->c.assign(v, v + p->n); break;
    case RODEO_FIELD_WGT_STATE: p->Q.assign(v, v + (size_t)p->n * p->n); break;
    case RODEO_FIELD_VAR_STATE: p->R.assign(v, v + (size_t)p->n * p->n); break;
    default: return fail(RODEO_E_INVALID, "unknown field");
  }
  p->tab_valid = false;          // the covariance tables depend on Q, R, W
  int rc = select_kernel_set(p); // the structure (hence the kernel family) may have changed
  if (rc == RODEO_OK && p->shared && !p->ks->sc_forward) p->shared = false;
  // (a compact variance layout stays selected: if the new prior has no kernel for it, the next solve fails with
  // RODEO_E_UNSUPPORTED instead of silently returning another layout)
  return rc;
}

// doubles per (system, time index) of var_out in the current layout
static int var_width(const rodeo_problem *p) {
  const int n = p->n, m = p->m;
  return p->var_layout == RODEO_VAR_BLOCKS ? m * (n / m) * (n / m) : p->var_layout == RODEO_VAR_DIAG ? n : n * n;
}

// dense wide-state kernels (kalman_dw.cuh) serve the plain solves of the sets that carry them
static bool wide_path(const rodeo_problem *p) {
  static const bool off = getenv("RODEO_CUDA_NO_WIDE") != nullptr;   // A/B against the rolled kernels
  return !off && p->ks->wide_hist_elems > 0 && !p->shared && p->ks->wide_forward[p->ig] != nullptr;
}

int rodeo_cuda_problem_info(const rodeo_problem *p, int32_t *blockdiag, int32_t *registers,
                            int32_t *hist_doubles_per_step) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (blockdiag) *blockdiag = p->ks->blockdiag;
  if (registers) *registers = p->ks->registers ? 1 : (wide_path(p) ? 2 : 0);
  if (hist_doubles_per_step) *hist_doubles_per_step = p->ks->hist_elems;
  return RODEO_OK;
}

// checkpoint interval actually used by entry point `bk`: the requested one where a kernel exists for it
static int effective_ck(const rodeo_problem *p, int bk = BK_MV) {
  // auto: interval 2 for the block-diagonal family (measured on B200: FN solve_mv 3.28 -> 2.85 ms,
  // log-likelihood sweep 2.88 -> 2.30 ms; interval 4 is slower again), 1 for the dense family
  // (its filter step is FP64-bound: re-running it took the backward kernel from 3.4 to 5.9 ms).
  // Families with >= 3 blocks have checkpointed kernels for solve_sim only (kalman_tpx.cuh).
  const int want = p->ck == 0 ? (p->ks->blockdiag ? 2 : 1) : p->ck;
  if (want == 1 || p->ig != RODEO_INTERROGATE_PROBDE) return 1;
  const int slot = ck_slot(want);
  return (slot >= 0 && p->ks->backward[slot][bk]) ? want : 1;
}

size_t rodeo_cuda_history_bytes_per_system(const rodeo_problem *p) {
  if (!p) return 0;
  if (p->shared) return (size_t)p->N * p->n * sizeof(double);   // filtered means only
  const size_t thread = (size_t)n_records(p->N, effective_ck(p)) * p->ks->hist_elems * sizeof(double);
  // the wide kernels keep (n + 1) rows of 16 lanes per step (padding lanes included)
  const size_t wide = wide_path(p) ? (size_t)p->N * p->ks->wide_hist_elems * sizeof(double) : 0;
  return wide > thread ? wide : thread;
}

int rodeo_cuda_set_shared_cov(rodeo_problem *p, int enabled) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (enabled && (!p->ks->sc_forward || p->ig == RODEO_INTERROGATE_CHKREBTII))
    return fail(RODEO_E_UNSUPPORTED,
                "shared-covariance mode needs a register kernel family and the probde or schobert interrogation");
  p->shared = enabled != 0;
  return RODEO_OK;
}

int rodeo_cuda_get_shared_cov(const rodeo_problem *p) { return p ? (p->shared ? 1 : 0) : RODEO_E_INVALID; }
int rodeo_cuda_problem_structured(const rodeo_problem *p) { return p ? (p->ks->structured ? 1 : 0) : RODEO_E_INVALID; }

int rodeo_cuda_set_var_layout(rodeo_problem *p, int layout) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (layout != RODEO_VAR_FULL && layout != RODEO_VAR_BLOCKS && layout != RODEO_VAR_DIAG)
    return fail(RODEO_E_INVALID, "unknown variance layout");
  if (layout != RODEO_VAR_FULL && (!p->ks->backward_vl[layout - 1][0] ||
                                   (p->ks->vl_ck > 1 && p->ig != RODEO_INTERROGATE_PROBDE)))
    return fail(RODEO_E_UNSUPPORTED,
                "compact variance layouts need a block-diagonal prior with at least two blocks");
  p->var_layout = layout;
  return RODEO_OK;
}
int rodeo_cuda_get_var_layout(const rodeo_problem *p) { return p ? p->var_layout : RODEO_E_INVALID; }
int rodeo_cuda_var_doubles_per_step(const rodeo_problem *p) { return p ? var_width(p) : RODEO_E_INVALID; }

int rodeo_cuda_set_checkpoint(rodeo_problem *p, int interval) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (interval != 0 && ck_slot(interval) < 0)
    return fail(RODEO_E_INVALID, "checkpoint interval must be 0 (auto), 1, 2 or 4");
  p->ck = interval;
  return RODEO_OK;
}

int rodeo_cuda_get_checkpoint(const rodeo_problem *p) { return p ? effective_ck(p) : RODEO_E_INVALID; }

int rodeo_cuda_set_workspace_limit(rodeo_problem *p, size_t bytes) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  p->ws_limit = bytes;
  return RODEO_OK;
}

}  // extern "C"

static long round_up(long x, long q) { return (x + q - 1) / q * q; }

// systems per chunk such that the history fits the workspace limit
static long chunk_capacity(const rodeo_problem *p, long batch) {
  size_t per = rodeo_cuda_history_bytes_per_system(p);
  long cap = (long)(p->ws_limit / per);
  cap = cap / 1024 * 1024;                 // whole CTAs, 256 B aligned history rows
  if (cap < 1024) cap = 1024;
  return std::min(cap, round_up(batch, 256));
}

static int ensure_workspace_bytes(rodeo_problem *p, size_t need);
static int ensure_workspace(rodeo_problem *p, long stride) {
  return ensure_workspace_bytes(p, rodeo_cuda_history_bytes_per_system(p) * (size_t)stride);
}
static int ensure_workspace_bytes(rodeo_problem *p, size_t need) {
  if (need <= p->ws_bytes) return RODEO_OK;
  if (p->ws) {
    cudaFree(p->ws);
    p->ws = nullptr;
    p->ws_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&p->ws, need);
  if (e != cudaSuccess) {
    cudaGetLastError();
    char buf[160];
    snprintf(buf, sizeof buf, "history workspace of %.2f GiB could not be allocated: %s",
             need / 1073741824.0, cudaGetErrorString(e));
    return fail(RODEO_E_NOMEM, buf);
  }
  p->ws_bytes = need;
  return RODEO_OK;
}

static int ensure_buf(rodeo_problem *p, int slot, size_t need) {
  if (need <= p->hbuf_bytes[slot]) return RODEO_OK;
  if (p->hbuf[slot]) cudaFree(p->hbuf[slot]);
  p->hbuf[slot] = nullptr;
  p->hbuf_bytes[slot] = 0;
  cudaError_t e = cudaMalloc(&p->hbuf[slot], need);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(RODEO_E_NOMEM, std::string("device staging buffer: ") + cudaGetErrorString(e));
  }
  p->hbuf_bytes[slot] = need;
  return RODEO_OK;
}

struct LoglikSpec {
  const int *thin_index = nullptr, *state_index = nullptr;
  const double *Y = nullptr;
  double gamma = 1.0;
  int n_obs_times = 0, n_obs_comp = 0;
  double *out = nullptr;
};

// Shared-covariance mode: (re)build the gain / Cholesky / variance tables for the current prior.
static int ensure_tables(rodeo_problem *p, const HostPrior &hp, cudaStream_t stream) {
  if (p->tab_valid) return RODEO_OK;
  const size_t need = (size_t)p->ks->sc_table_doubles(p->N) * sizeof(double);
  if (need > p->tab_bytes) {
    if (p->tab) cudaFree(p->tab);
    p->tab = nullptr;
    p->tab_bytes = 0;
    if (cudaMalloc((void **)&p->tab, need) != cudaSuccess) {
      cudaGetLastError();
      return fail(RODEO_E_NOMEM, "shared-covariance tables could not be allocated");
    }
    p->tab_bytes = need;
  }
  if (!p->tab_fail_dev) CK(cudaMalloc((void **)&p->tab_fail_dev, sizeof(int)));
  CK(cudaMemsetAsync(p->tab, 0, need, stream));
  CK(cudaMemsetAsync(p->tab_fail_dev, 0, sizeof(int), stream));
  CK(p->ks->sc_tables(hp, p->ig == RODEO_INTERROGATE_SCHOBERT, p->tab, p->tab_fail_dev, stream));
  CK(cudaMemcpyAsync(&p->tab_fail, p->tab_fail_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));   // one-off per prior: the failure flag travels in SolveArgs
  p->launches += 1;
  p->tab_valid = true;
  return RODEO_OK;
}

// A/B knob for measurements: RODEO_CUDA_BACKWARD=staged selects the round-1 smoother (one thread per system,
// one run per visit) where the thread-per-block kernel (kalman_tpb.cuh) replaced it.  Read once per process.
static bool prefer_staged_backward() {
  static const bool v = [] {
    const char *e = getenv("RODEO_CUDA_BACKWARD");
    return e && std::string(e) == "staged";
  }();
  return v;
}

// Per-system priors: is EVERY system's prior block diagonal with m blocks of size n / m and row a of W inside block a?
// One thread per system; flag = 1 on the first violation.
__global__ void check_batched_blockdiag(BatchedPrior bp, long B, int n, int m, int *flag) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int bs = n / m;
  bool bad = false;
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++)
      if (i / bs != j / bs) {
        if (bp.Q && bp.Q[b * n * n + i + (long)j * n] != 0.0) bad = true;
        if (bp.R && bp.R[b * n * n + i + (long)j * n] != 0.0) bad = true;
      }
  if (bp.W)
    for (int j = 0; j < n; j++)
      for (int a = 0; a < m; a++)
        if (j / bs != a && bp.W[b * m * n + a + (long)j * m] != 0.0) bad = true;
  if (bad) *flag = 1;
}

// One pass over `batch` systems in chunks: forward kernel then backward kernel per chunk.
static int run_chunks(rodeo_problem *p, BackwardKind bk, long batch, const double *x0,
                      const double *theta, const double *z, long z_stride, double *x_out,
                      double *mu_out, double *var_out, int *status, const LoglikSpec *ll,
                      cudaStream_t stream, const BatchedPrior *bp = nullptr, bool bp_bd = false) {
  const int n = p->n, nth = ode_n_theta(p->ode);
  const long T1 = (long)p->N + 1;
  // per-system priors: block-diagonal thread-per-block kernels (checkpoint 2, permuted history) when every
  // system's prior has the structure (bp_bd), else the dense family with the full history; never shared-covariance
  const KernelSet *ks = (bp && !bp_bd) ? dense_set(p) : p->ks;
  const bool shared = p->shared && !bp;
  const bool wide = !bp && wide_path(p) && p->ks->wide_backward[bk] != nullptr;
  // a compact variance layout runs on its own kernels (checkpoint 2) for solve_mv / solve
  const bool vl = p->var_layout != RODEO_VAR_FULL && !bp && (bk == BK_MV || bk == BK_BOTH);
  if (vl && (p->shared || !p->ks->backward_vl[p->var_layout - 1][0]))
    return fail(RODEO_E_UNSUPPORTED, "the variance layout is not available in this mode (shared covariance / kernel family)");
  const int ck = vl ? p->ks->vl_ck : bp ? (bp_bd ? 2 : 1) : effective_ck(p, bk), slot = ck_slot(ck);
  const int vw = (bp || p->shared) ? n * n : var_width(p);
  const bool wide_pp = bp && !bp_bd && ks->wide_pp_forward != nullptr;   // dense wide-state kernels, per-system priors
  // history bytes per system of THIS entry point (rodeo_cuda_history_bytes_per_system is the largest over them)
  const size_t per_system = bp_bd ? (size_t)n_records(p->N, ck) * ks->hist_elems * sizeof(double)
                            : bp ? (size_t)p->N * (wide_pp ? ks->wide_hist_elems : ks->hist_elems) * sizeof(double)
                               : (shared || wide) ? rodeo_cuda_history_bytes_per_system(p)
                                                  : (size_t)n_records(p->N, ck) * ks->hist_elems * sizeof(double);
  long cap = (long)(p->ws_limit / per_system) / 1024 * 1024;
  if (cap < 1024) cap = 1024;
  cap = std::min(cap, round_up(batch, 256));   // whole 256-system blocks: the permuted history order pads to them
  int rc = ensure_workspace_bytes(p, per_system * (size_t)cap);
  if (rc) return rc;
  HostPrior hp{p->n, p->m, p->N, p->tmin, p->tmax, p->W.data(), p->Q.data(), p->c.data(), p->R.data()};
  if (shared) {
    if ((rc = ensure_tables(p, hp, stream))) return rc;
    if (var_out)   // ONE (N+1, n, n) array serves the whole batch in this mode
      CK(cudaMemcpyAsync(var_out, p->tab + ks->sc_table_var_offset(p->N), (size_t)T1 * n * n * sizeof(double),
                         cudaMemcpyDeviceToDevice, stream));
  }
  // which smoother runs: the current kernel, or the previous generation on request (A/B measurements)
  launch_fn backward = nullptr;
  bool perm = bp_bd;
  // the thread-per-block kernels keep 32-bit byte offsets inside a 16-system output row: trajectories longer than
  // that (16 * (N+1) * n^2 * 8 >= 2^31, i.e. ~466,000 steps for n_state 6) run on the previous generation
  const bool long_rows = 16.0 * (double)T1 * n * n * 8.0 >= p->row_limit;
  if (vl) {
    if (long_rows && ks->vl_perm)
      return fail(RODEO_E_UNSUPPORTED, "compact variance layouts: n_eval is too large for the thread-per-block kernels");
    backward = ks->backward_vl[p->var_layout - 1][bk == BK_BOTH ? 1 : 0];
    perm = ks->vl_perm;
  } else if (!bp && !shared && !wide) {
    const bool alt = (prefer_staged_backward() || (long_rows && ks->backward_perm[slot][bk])) && ks->backward_alt[slot][bk];
    backward = alt ? ks->backward_alt[slot][bk] : ks->backward[slot][bk];
    perm = !alt && ks->backward_perm[slot][bk];
  }
  const long n_chunks = (batch + cap - 1) / cap;
  const int kMaxTimedChunks = 4096;       // beyond that only the last solve is kept (the ring restarts)
  if (p->timing) {
    if (p->ev_chunks + n_chunks > kMaxTimedChunks) p->ev_chunks = p->ev_solves = 0;
    while ((long)p->ev.size() < 3 * (p->ev_chunks + n_chunks)) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      p->ev.push_back(e);
    }
    p->ev_first = p->ev_chunks;
    p->ev_chunks += (int)n_chunks;
    p->ev_solves += 1;
  }
  const long ev0 = p->timing ? p->ev_first : 0;
  for (long ci = 0; ci < n_chunks; ci++) {
    const long off = ci * cap, B = std::min(cap, batch - off);
    SolveArgs a;
    memset(&a, 0, sizeof a);
    a.B = B;
    a.x0 = x0 + off * n;
    a.theta = theta ? theta + off * nth : nullptr;
    a.z = z ? z + off * z_stride : nullptr;
    a.z_stride = z_stride;
    a.hist = (double *)p->ws;
    a.ck = ck;
    a.perm16 = perm ? 1 : 0;
    a.n_steps = p->N + 1;
    a.x_out = x_out ? x_out + off * T1 * n : nullptr;
    a.mu_out = mu_out ? mu_out + off * T1 * n : nullptr;
    a.var_out = (var_out && !shared) ? var_out + off * T1 * vw : nullptr;
    a.tab = shared ? p->tab : nullptr;
    a.tab_fail = shared ? p->tab_fail : 0;
    a.status = status ? status + off : nullptr;
    if (ll) {
      a.thin_index = ll->thin_index;
      a.state_index = ll->state_index;
      a.Y = ll->Y;
      a.gamma = ll->gamma;
      a.n_obs_times = ll->n_obs_times;
      a.n_obs_comp = ll->n_obs_comp;
      a.loglik = ll->out + off;
    }
    if ((reinterpret_cast<uintptr_t>(a.x_out) | reinterpret_cast<uintptr_t>(a.mu_out) |
         reinterpret_cast<uintptr_t>(a.var_out)) & 15)
      return fail(RODEO_E_INVALID, "internal: a chunk's output array is not 16-byte aligned");
    if (p->timing) CK(cudaEventRecord(p->ev[3 * (ev0 + ci) + 0], stream));
    if (bp) {
      BatchedPrior c{bp->W ? bp->W + off * p->m * n : nullptr, bp->Q ? bp->Q + off * n * n : nullptr,
                     bp->c ? bp->c + off * n : nullptr, bp->R ? bp->R + off * n * n : nullptr};
      CK(bp_bd ? ks->pp_bd_forward(hp, c, a, stream)
               : wide_pp ? ks->wide_pp_forward(hp, c, a, stream) : ks->pp_forward(hp, c, a, stream));
      if (p->timing) CK(cudaEventRecord(p->ev[3 * (ev0 + ci) + 1], stream));
      CK(bp_bd ? ks->pp_bd_backward[bk](hp, c, a, stream)
               : wide_pp ? ks->wide_pp_backward[bk](hp, c, a, stream) : ks->pp_backward[bk](hp, c, a, stream));
    } else {
      CK(shared ? ks->sc_forward(hp, a, stream)
                : wide ? ks->wide_forward[p->ig](hp, a, stream)
                       : (perm && ks->forward_perm[p->ig]) ? ks->forward_perm[p->ig](hp, a, stream)
                                                           : ks->forward[p->ig](hp, a, stream));
      if (p->timing) CK(cudaEventRecord(p->ev[3 * (ev0 + ci) + 1], stream));
      CK(shared ? ks->sc_backward[bk](hp, a, stream)
                : wide ? ks->wide_backward[bk](hp, a, stream) : backward(hp, a, stream));
    }
    if (p->timing) CK(cudaEventRecord(p->ev[3 * (ev0 + ci) + 2], stream));
    p->launches += 2;
  }
  return RODEO_OK;
}

static int check_solve_args(rodeo_problem *p, int mode, int64_t batch, const double *x0,
                            const double *theta, const double *z, double *x_out, double *mu_out,
                            double *var_out) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (mode != RODEO_MODE_MV && mode != RODEO_MODE_SIM && mode != RODEO_MODE_BOTH)
    return fail(RODEO_E_INVALID, "mode must be RODEO_MODE_MV, _SIM or _BOTH");
  if (batch < 0) return fail(RODEO_E_INVALID, "negative batch");
  if (batch > 0 && !x0) return fail(RODEO_E_INVALID, "x0 is null");
  if (batch > 0 && ode_n_theta(p->ode) > 0 && !theta) return fail(RODEO_E_INVALID, "theta is null but the RHS has parameters");
  if (batch > 0 && ((mode & RODEO_MODE_SIM) || p->ig == RODEO_INTERROGATE_CHKREBTII) && !z)
    return fail(RODEO_E_INVALID, "z is null but the requested mode consumes normal draws");
  if (batch > 0 && (mode & RODEO_MODE_MV) && (!mu_out || !var_out)) return fail(RODEO_E_INVALID, "mu_out / var_out is null");
  if (batch > 0 && (mode & RODEO_MODE_SIM) && !x_out) return fail(RODEO_E_INVALID, "x_out is null");
  return RODEO_OK;
}

// Device outputs are written with 16-byte vector / TMA stores.
static int check_device_alignment(const double *x_out, const double *mu_out, const double *var_out) {
  if ((reinterpret_cast<uintptr_t>(x_out) | reinterpret_cast<uintptr_t>(mu_out) | reinterpret_cast<uintptr_t>(var_out)) & 15)
    return fail(RODEO_E_INVALID, "device output arrays must be 16-byte aligned");
  return RODEO_OK;
}

// Can the per-system priors of this call run on the block-diagonal thread-per-block kernels?  The handle's own
// prior must have selected a block-diagonal set that carries them, and every system's arrays must have the
// structure (checked on the device; one 4-byte read back -- this entry point synchronises the stream once).
static int batched_prior_is_blockdiag(rodeo_problem *p, const BatchedPrior &bp, long batch, cudaStream_t stream,
                                      bool *out) {
  *out = false;
  if (!p->ks->blockdiag || !p->ks->pp_bd_forward || (p->flags & RODEO_FLAG_FORCE_DENSE)) return RODEO_OK;
  if (!p->bd_flag_dev) CK(cudaMalloc((void **)&p->bd_flag_dev, sizeof(int)));
  CK(cudaMemsetAsync(p->bd_flag_dev, 0, sizeof(int), stream));
  check_batched_blockdiag<<<(unsigned)((batch + 127) / 128), 128, 0, stream>>>(bp, batch, p->n, p->m, p->bd_flag_dev);
  CK(cudaGetLastError());
  int flag = 1;
  CK(cudaMemcpyAsync(&flag, p->bd_flag_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  *out = flag == 0;
  return RODEO_OK;
}

extern "C" {

int rodeo_cuda_reserve(rodeo_problem *p, int64_t max_batch) {
  if (!p || max_batch < 0) return fail(RODEO_E_INVALID, "bad arguments");
  if (max_batch == 0) return RODEO_OK;
  CK(cudaSetDevice(p->device));
  return ensure_workspace(p, chunk_capacity(p, (long)max_batch));
}

int rodeo_cuda_solve(rodeo_problem *p, int mode, int64_t batch, const double *x0, const double *theta,
                     const double *z, int64_t z_batch_stride, double *x_out, double *mu_out,
                     double *var_out, int32_t *status, void *cuda_stream) {
  int rc = check_solve_args(p, mode, batch, x0, theta, z, x_out, mu_out, var_out);
  if (!rc && batch > 0) rc = check_device_alignment(x_out, mu_out, var_out);
  if (rc) return rc;
  if (batch == 0) return RODEO_OK;
  CK(cudaSetDevice(p->device));
  BackwardKind bk = mode == RODEO_MODE_MV ? BK_MV : mode == RODEO_MODE_SIM ? BK_SIM : BK_BOTH;
  return run_chunks(p, bk, (long)batch, x0, theta, z, (long)z_batch_stride, x_out, mu_out, var_out,
                    status, nullptr, (cudaStream_t)cuda_stream);
}

int rodeo_cuda_solve_batched_prior(rodeo_problem *p, int mode, int64_t batch, const double *x0,
                                   const double *theta, const double *z, int64_t z_batch_stride,
                                   const double *wgt_meas_b, const double *mu_state_b,
                                   const double *wgt_state_b, const double *var_state_b, double *x_out,
                                   double *mu_out, double *var_out, int32_t *status, void *cuda_stream) {
  int rc = check_solve_args(p, mode, batch, x0, theta, z, x_out, mu_out, var_out);
  if (!rc && batch > 0) rc = check_device_alignment(x_out, mu_out, var_out);
  if (rc) return rc;
  if (batch == 0) return RODEO_OK;
  if (p->ig != RODEO_INTERROGATE_PROBDE)
    return fail(RODEO_E_UNSUPPORTED, "per-system priors need the probde interrogation");
  CK(cudaSetDevice(p->device));
  BatchedPrior bp{wgt_meas_b, wgt_state_b, mu_state_b, var_state_b};
  BackwardKind bk = mode == RODEO_MODE_MV ? BK_MV : mode == RODEO_MODE_SIM ? BK_SIM : BK_BOTH;
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  bool bd = false;
  rc = batched_prior_is_blockdiag(p, bp, (long)batch, stream, &bd);
  if (rc) return rc;
  if (16.0 * (double)(p->N + 1) * p->n * p->n * 8.0 >= p->row_limit) bd = false;   // 32-bit row offsets (see run_chunks)
  p->last_pp_bd = bd;
  if (!bd) {
    const KernelSet *dense = dense_set(p);
    if (!dense || !dense->pp_forward)
      return fail(RODEO_E_UNSUPPORTED, "per-system priors that are not block diagonal need a dense kernel set");
  }
  return run_chunks(p, bk, (long)batch, x0, theta, z, (long)z_batch_stride, x_out, mu_out, var_out, status,
                    nullptr, stream, &bp, bd);
}

int rodeo_cuda_last_batched_prior_blockdiag(const rodeo_problem *p) { return p ? (p->last_pp_bd ? 1 : 0) : RODEO_E_INVALID; }

int rodeo_cuda_loglik_batched_prior(rodeo_problem *p, int use_draw, int64_t batch, const double *x0,
                                    const double *theta, const double *z, int64_t z_batch_stride,
                                    const double *wgt_meas_b, const double *mu_state_b,
                                    const double *wgt_state_b, const double *var_state_b,
                                    const int32_t *thin_index, int32_t n_obs_times, const int32_t *state_index,
                                    int32_t n_obs_comp, const double *Y, double gamma, double *loglik_out,
                                    int32_t *status, void *cuda_stream) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (batch < 0 || n_obs_times < 0 || n_obs_comp < 0) return fail(RODEO_E_INVALID, "negative size");
  if (batch == 0) return RODEO_OK;
  if (!x0 || !loglik_out) return fail(RODEO_E_INVALID, "x0 / loglik_out is null");
  if (ode_n_theta(p->ode) > 0 && !theta) return fail(RODEO_E_INVALID, "theta is null but the RHS has parameters");
  if (use_draw && !z) return fail(RODEO_E_INVALID, "z is null but draws are requested");
  if (n_obs_times * n_obs_comp > 0 && (!thin_index || !state_index || !Y)) return fail(RODEO_E_INVALID, "thin_index / state_index / Y is null");
  if (!(gamma > 0.0)) return fail(RODEO_E_INVALID, "gamma must be positive");
  if (p->ig != RODEO_INTERROGATE_PROBDE)
    return fail(RODEO_E_UNSUPPORTED, "per-system priors need the probde interrogation");
  const KernelSet *dense = dense_set(p);
  const BackwardKind bk = use_draw ? BK_LL_DRAW : BK_LL_MEAN;
  if (!dense || !dense->pp_forward || !(dense->wide_pp_forward ? dense->wide_pp_backward[bk] : dense->pp_backward[bk]))
    return fail(RODEO_E_UNSUPPORTED, "log-likelihood with per-system priors needs a dense kernel set");
  CK(cudaSetDevice(p->device));
  BatchedPrior bp{wgt_meas_b, wgt_state_b, mu_state_b, var_state_b};
  LoglikSpec ll;
  ll.thin_index = thin_index;
  ll.state_index = state_index;
  ll.Y = Y;
  ll.gamma = gamma;
  ll.n_obs_times = n_obs_times;
  ll.n_obs_comp = n_obs_comp;
  ll.out = loglik_out;
  return run_chunks(p, bk, (long)batch, x0, theta, z, (long)z_batch_stride, nullptr, nullptr, nullptr, status, &ll,
                    (cudaStream_t)cuda_stream, &bp, false);
}

int rodeo_cuda_loglik(rodeo_problem *p, int use_draw, int64_t batch, const double *x0, const double *theta,
                      const double *z, int64_t z_batch_stride, const int32_t *thin_index,
                      int32_t n_obs_times, const int32_t *state_index, int32_t n_obs_comp,
                      const double *Y, double gamma, double *loglik_out, int32_t *status,
                      void *cuda_stream) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (batch < 0 || n_obs_times < 0 || n_obs_comp < 0) return fail(RODEO_E_INVALID, "negative size");
  if (batch == 0) return RODEO_OK;
  if (!x0 || !loglik_out) return fail(RODEO_E_INVALID, "x0 / loglik_out is null");
  if (ode_n_theta(p->ode) > 0 && !theta) return fail(RODEO_E_INVALID, "theta is null but the RHS has parameters");
  if ((use_draw || p->ig == RODEO_INTERROGATE_CHKREBTII) && !z) return fail(RODEO_E_INVALID, "z is null but draws are requested");
  if (n_obs_times * n_obs_comp > 0 && (!thin_index || !state_index || !Y)) return fail(RODEO_E_INVALID, "thin_index / state_index / Y is null");
  if (!(gamma > 0.0)) return fail(RODEO_E_INVALID, "gamma must be positive");
  CK(cudaSetDevice(p->device));
  LoglikSpec ll;
  ll.thin_index = thin_index;
  ll.state_index = state_index;
  ll.Y = Y;
  ll.gamma = gamma;
  ll.n_obs_times = n_obs_times;
  ll.n_obs_comp = n_obs_comp;
  ll.out = loglik_out;
  return run_chunks(p, use_draw ? BK_LL_DRAW : BK_LL_MEAN, (long)batch, x0, theta, z, (long)z_batch_stride,
                    nullptr, nullptr, nullptr, status, &ll, (cudaStream_t)cuda_stream);
}

// Host-buffer path: H2D inputs, solve chunk k on the compute stream into one of two device
// output buffers while the copy stream drains chunk k-1 to the host.
int rodeo_cuda_solve_host(rodeo_problem *p, int mode, int64_t batch, const double *x0, const double *theta,
                          const double *z, int64_t z_batch_stride, double *x_out, double *mu_out,
                          double *var_out, int32_t *status) {
  int rc = check_solve_args(p, mode, batch, x0, theta, z, x_out, mu_out, var_out);
  if (rc) return rc;
  if (batch == 0) return RODEO_OK;
  CK(cudaSetDevice(p->device));
  const int n = p->n, nth = ode_n_theta(p->ode);
  const long T1 = (long)p->N + 1;
  if (!p->s_compute) {
    CK(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&p->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CK(cudaEventCreateWithFlags(&p->e_done[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&p->e_copied[i], cudaEventDisableTiming));
    }
  }
  const bool mv = mode & RODEO_MODE_MV, sim = mode & RODEO_MODE_SIM;
  const bool pervar = mv && !p->shared;   // per-system variance output (general mode)
  const int vw = var_width(p);                // doubles of variance per (system, time index) in the current layout
  const size_t out_per = (size_t)T1 * sizeof(double) * ((mv ? n : 0) + (pervar ? vw : 0) + (sim ? n : 0));
  // sub-arrays of a staging buffer start on 128-byte boundaries (the kernels store 16-byte vectors / TMA
  // boxes relative to each array base): B * T1 * n doubles is odd when B, T1 and n are all odd
  auto al = [](size_t doubles) { return (doubles + 15) & ~(size_t)15; };
  const size_t out_pad = 3 * 16 * sizeof(double);
  // output staging: two buffers of at most 2 GiB each, and never more systems than a history chunk
  long hc = (long)std::max<size_t>(1, ((size_t)2 << 30) / out_per);
  hc = std::min(hc, chunk_capacity(p, (long)batch));
  hc = std::min(hc, (long)batch);
  if (hc > 1024) hc = hc / 1024 * 1024;
  const bool need_z = z != nullptr;
  const size_t z_elems = need_z ? (size_t)n * p->N : 0;
  // slots: 0 x0, 1 theta, 2 z, 3 status, 4/5 output buffers
  if ((rc = ensure_buf(p, 0, (size_t)hc * n * sizeof(double)))) return rc;
  if (nth && (rc = ensure_buf(p, 1, (size_t)hc * nth * sizeof(double)))) return rc;
  if (need_z && (rc = ensure_buf(p, 2, (z_batch_stride ? (size_t)hc : 1) * z_elems * sizeof(double)))) return rc;
  if ((rc = ensure_buf(p, 3, (size_t)hc * sizeof(int)))) return rc;
  if ((rc = ensure_buf(p, 4, (size_t)hc * out_per + out_pad))) return rc;
  if ((rc = ensure_buf(p, 5, (size_t)hc * out_per + out_pad))) return rc;
  if (mv && p->shared && (rc = ensure_buf(p, 6, (size_t)T1 * n * n * sizeof(double)))) return rc;
  if (need_z && !z_batch_stride)
    CK(cudaMemcpyAsync(p->hbuf[2], z, z_elems * sizeof(double), cudaMemcpyHostToDevice, p->s_compute));
  BackwardKind bk = mode == RODEO_MODE_MV ? BK_MV : mode == RODEO_MODE_SIM ? BK_SIM : BK_BOTH;
  const long n_chunks = ((long)batch + hc - 1) / hc;
  for (long ci = 0; ci < n_chunks; ci++) {
    const long off = ci * hc, B = std::min(hc, (long)batch - off);
    const int slot = (int)(ci & 1);
    double *ob = (double *)p->hbuf[4 + slot];
    const size_t vec = al((size_t)B * T1 * n), mat = al((size_t)B * T1 * vw);
    double *d_mu = mv ? ob : nullptr;
    double *d_var = pervar ? ob + vec : nullptr;
    double *d_x = sim ? ob + (mv ? vec + (pervar ? mat : 0) : 0) : nullptr;
    if (mv && p->shared && ci == 0) d_var = (double *)p->hbuf[6];   // the single shared variance array
    // inputs of this chunk (the previous chunk's kernels are ordered before us on s_compute)
    CK(cudaMemcpyAsync(p->hbuf[0], x0 + off * n, (size_t)B * n * sizeof(double), cudaMemcpyHostToDevice, p->s_compute));
    if (nth) CK(cudaMemcpyAsync(p->hbuf[1], theta + off * nth, (size_t)B * nth * sizeof(double), cudaMemcpyHostToDevice, p->s_compute));
    if (need_z && z_batch_stride)
      CK(cudaMemcpyAsync(p->hbuf[2], z + off * z_batch_stride, (size_t)B * z_elems * sizeof(double), cudaMemcpyHostToDevice, p->s_compute));
    if (ci >= 2) CK(cudaStreamWaitEvent(p->s_compute, p->e_copied[slot], 0));
    rc = run_chunks(p, bk, B, (const double *)p->hbuf[0], nth ? (const double *)p->hbuf[1] : nullptr,
                    need_z ? (const double *)p->hbuf[2] : nullptr, z_batch_stride ? (long)z_elems : 0, d_x,
                    d_mu, d_var, (int *)p->hbuf[3], nullptr, p->s_compute);
    if (rc) return rc;
    if (status) CK(cudaMemcpyAsync(status + off, p->hbuf[3], (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, p->s_compute));
    CK(cudaEventRecord(p->e_done[slot], p->s_compute));
    CK(cudaStreamWaitEvent(p->s_copy, p->e_done[slot], 0));
    if (mv) {
      CK(cudaMemcpyAsync(mu_out + off * T1 * n, d_mu, (size_t)B * T1 * n * sizeof(double), cudaMemcpyDeviceToHost, p->s_copy));
      if (pervar)
        CK(cudaMemcpyAsync(var_out + off * T1 * vw, d_var, (size_t)B * T1 * vw * sizeof(double), cudaMemcpyDeviceToHost, p->s_copy));
      else if (ci == 0)
        CK(cudaMemcpyAsync(var_out, d_var, (size_t)T1 * n * n * sizeof(double), cudaMemcpyDeviceToHost, p->s_copy));
    }
    if (sim) CK(cudaMemcpyAsync(x_out + off * T1 * n, d_x, (size_t)B * T1 * n * sizeof(double), cudaMemcpyDeviceToHost, p->s_copy));
    CK(cudaEventRecord(p->e_copied[slot], p->s_copy));
  }
  CK(cudaStreamSynchronize(p->s_compute));
  CK(cudaStreamSynchronize(p->s_copy));
  return RODEO_OK;
}

// Debug / test hook: copy the first `bytes` of the history workspace to a device buffer.
int rodeo_cuda_debug_copy_history(rodeo_problem *p, void *dst_device, size_t bytes, void *cuda_stream) {
  if (!p || !dst_device) return fail(RODEO_E_INVALID, "null argument");
  if (bytes > p->ws_bytes) return fail(RODEO_E_INVALID, "more bytes than the workspace holds");
  CK(cudaMemcpyAsync(dst_device, p->ws, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
  return RODEO_OK;
}

// Test hook: lower the output-row size above which the thread-per-block kernels hand over to the previous
// generation (default 2^31 bytes), so that the hand-over can be exercised on ordinary problem sizes.
int rodeo_cuda_debug_set_row_limit(rodeo_problem *p, double bytes) {
  if (!p || !(bytes > 0.0) || bytes > 2147483648.0) return fail(RODEO_E_INVALID, "bad arguments");
  p->row_limit = bytes;
  return RODEO_OK;
}

int64_t rodeo_cuda_launch_count(const rodeo_problem *p) { return p ? p->launches : 0; }
int rodeo_cuda_reset_launch_count(rodeo_problem *p) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  p->launches = 0;
  return RODEO_OK;
}
int rodeo_cuda_set_timing(rodeo_problem *p, int enabled) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  p->timing = enabled != 0;
  p->ev_first = p->ev_chunks = p->ev_solves = 0;
  return RODEO_OK;
}
// forward / backward kernel time summed over chunks [c0, c1) of the recorded events
static int sum_kernel_ms(rodeo_problem *p, int c0, int c1, double *f, double *b) {
  CK(cudaSetDevice(p->device));
  *f = *b = 0.0;
  for (int ci = c0; ci < c1; ci++) {
    CK(cudaEventSynchronize(p->ev[3 * ci + 2]));
    float t;
    CK(cudaEventElapsedTime(&t, p->ev[3 * ci + 0], p->ev[3 * ci + 1]));
    *f += t;
    CK(cudaEventElapsedTime(&t, p->ev[3 * ci + 1], p->ev[3 * ci + 2]));
    *b += t;
  }
  return RODEO_OK;
}
int rodeo_cuda_last_kernel_ms(rodeo_problem *p, float *forward_ms, float *backward_ms) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (!p->timing || p->ev_chunks == 0) return fail(RODEO_E_INVALID, "timing was not enabled for the last solve");
  double f, b;
  int rc = sum_kernel_ms(p, p->ev_first, p->ev_chunks, &f, &b);
  if (rc) return rc;
  if (forward_ms) *forward_ms = (float)f;
  if (backward_ms) *backward_ms = (float)b;
  return RODEO_OK;
}
int rodeo_cuda_kernel_ms_total(rodeo_problem *p, double *forward_ms, double *backward_ms, int32_t *n_solves) {
  if (!p) return fail(RODEO_E_INVALID, "null problem");
  if (!p->timing || p->ev_chunks == 0) return fail(RODEO_E_INVALID, "no solve was timed since rodeo_cuda_set_timing");
  double f, b;
  int rc = sum_kernel_ms(p, 0, p->ev_chunks, &f, &b);
  if (rc) return rc;
  if (forward_ms) *forward_ms = f;
  if (backward_ms) *backward_ms = b;
  if (n_solves) *n_solves = p->ev_solves;
  return RODEO_OK;
}

}  // extern "C"
