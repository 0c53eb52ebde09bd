// registry.cu -- collects the compiled kernel sets.
#include "registry.cuh"

#include <mutex>

namespace rodeo {

#define RODEO_REG(name) void register_##name(KernelSet *, int *);
RODEO_REG(chkrebtii_dense) RODEO_REG(fitz_dense) RODEO_REG(lorenz_dense) RODEO_REG(mseir_dense)
RODEO_REG(mixed)
RODEO_REG(chkrebtii_bd_p3)
RODEO_REG(chkrebtii_bd_p4)
RODEO_REG(chkrebtii_bds_p3)
RODEO_REG(chkrebtii_bds_p4)
RODEO_REG(fitz_bd_p2)
RODEO_REG(fitz_bd_p3)
RODEO_REG(fitz_bd_p4)
RODEO_REG(fitz_bds_p2)
RODEO_REG(fitz_bds_p3)
RODEO_REG(fitz_bds_p4)
RODEO_REG(lorenz_bd_p2)
RODEO_REG(lorenz_bd_p3)
RODEO_REG(lorenz_bd_p4)
RODEO_REG(lorenz_bds_p2)
RODEO_REG(lorenz_bds_p3)
RODEO_REG(lorenz_bds_p4)
RODEO_REG(mseir_bd_p2)
RODEO_REG(mseir_bd_p3)
RODEO_REG(mseir_bds_p2)
RODEO_REG(mseir_bds_p3)
#undef RODEO_REG

void attach_wide(KernelSet *, int);   // kernels_wide.cu

static KernelSet g_sets[64];
static int g_count = -1;
static std::once_flag g_once;

static void fill_sets() {
  int c = 0;
  register_chkrebtii_dense(g_sets, &c);
  register_fitz_dense(g_sets, &c);
  register_lorenz_dense(g_sets, &c);
  register_mseir_dense(g_sets, &c);
  register_mixed(g_sets, &c);
  register_chkrebtii_bd_p3(g_sets, &c);
  register_chkrebtii_bd_p4(g_sets, &c);
  register_chkrebtii_bds_p3(g_sets, &c);
  register_chkrebtii_bds_p4(g_sets, &c);
  register_fitz_bd_p2(g_sets, &c);
  register_fitz_bd_p3(g_sets, &c);
  register_fitz_bd_p4(g_sets, &c);
  register_fitz_bds_p2(g_sets, &c);
  register_fitz_bds_p3(g_sets, &c);
  register_fitz_bds_p4(g_sets, &c);
  register_lorenz_bd_p2(g_sets, &c);
  register_lorenz_bd_p3(g_sets, &c);
  register_lorenz_bd_p4(g_sets, &c);
  register_lorenz_bds_p2(g_sets, &c);
  register_lorenz_bds_p3(g_sets, &c);
  register_lorenz_bds_p4(g_sets, &c);
  register_mseir_bd_p2(g_sets, &c);
  register_mseir_bd_p3(g_sets, &c);
  register_mseir_bds_p2(g_sets, &c);
  register_mseir_bds_p3(g_sets, &c);
  attach_wide(g_sets, c);
  // A structured set specialises only its FP64-bound kernels (forward filters, log-likelihood and draw-only
  // smoothers).  Its HBM-bound, output-writing smoothers, the compact-layout and per-system-prior launchers ARE the
  // general twin's (bit-identical results, nothing compiled twice): fill the empty slots from it.
  for (int i = 0; i < c; i++) {
    KernelSet &k = g_sets[i];
    if (!k.structured) continue;
    const KernelSet *g = nullptr;
    for (int j = 0; j < c; j++)
      if (!g_sets[j].structured && g_sets[j].blockdiag && g_sets[j].ode == k.ode && g_sets[j].n == k.n) g = &g_sets[j];
    if (!g) continue;
    for (int ci = 0; ci < kNumCk; ci++)
      for (int b = 0; b < BK_COUNT; b++) {
        if (!k.backward[ci][b] && g->backward[ci][b]) {
          k.backward[ci][b] = g->backward[ci][b];
          k.backward_perm[ci][b] = g->backward_perm[ci][b];
        }
        if (!k.backward_alt[ci][b]) k.backward_alt[ci][b] = g->backward_alt[ci][b];
      }
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 2; b++)
        if (!k.backward_vl[a][b]) {
          k.backward_vl[a][b] = g->backward_vl[a][b];
          k.vl_ck = g->vl_ck;
          k.vl_perm = g->vl_perm;
        }
    if (!k.pp_bd_forward) {
      k.pp_bd_forward = g->pp_bd_forward;
      for (int b = 0; b < BK_COUNT; b++) k.pp_bd_backward[b] = g->pp_bd_backward[b];
    }
  }
  g_count = c;
}
static void init_sets() { std::call_once(g_once, fill_sets); }   // handles may be created from several threads

// n_meas / n_theta of a registered right-hand side, taken from its functor (ode_functors.cuh)
int ode_count() { return 4; }
int ode_n_meas(int ode) {
  switch (ode) {
    case OdeChkrebtii::id: return OdeChkrebtii::n_meas;
    case OdeFitz::id: return OdeFitz::n_meas;
    case OdeLorenz63::id: return OdeLorenz63::n_meas;
    case OdeMseir::id: return OdeMseir::n_meas;
    default: return -1;
  }
}
int ode_order(int ode) {
  switch (ode) {
    case OdeChkrebtii::id: return OdeChkrebtii::order;
    case OdeFitz::id: return OdeFitz::order;
    case OdeLorenz63::id: return OdeLorenz63::order;
    case OdeMseir::id: return OdeMseir::order;
    default: return -1;
  }
}
int ode_n_theta(int ode) {
  switch (ode) {
    case OdeChkrebtii::id: return OdeChkrebtii::n_theta;
    case OdeFitz::id: return OdeFitz::n_theta;
    case OdeLorenz63::id: return OdeLorenz63::n_theta;
    case OdeMseir::id: return OdeMseir::n_theta;
    default: return -1;
  }
}

const KernelSet *find_kernel_set(int ode, int n, bool blockdiag, bool structured, const int *off) {
  init_sets();
  for (int i = 0; i < g_count; i++) {
    const KernelSet &k = g_sets[i];
    if (k.ode != ode || k.n != n || k.blockdiag != blockdiag || k.structured != structured) continue;
    if (!off) {
      if (!k.mixed) return &k;
      continue;
    }
    bool same = true;
    for (int a = 0; a < k.m; a++) same = same && k.off[a] == off[a];
    if (same) return &k;
  }
  return nullptr;
}
int kernel_set_count() {
  init_sets();
  return g_count;
}
const KernelSet *kernel_set_at(int i) {
  init_sets();
  return (i >= 0 && i < g_count) ? &g_sets[i] : nullptr;
}

}  // namespace rodeo
