// registry.cu -- collects the compiled kernel sets.
#include "registry.cuh"

#include <mutex>

namespace rodeo {

#define RODEO_REG(name) void register_##name(KernelSet *, int *);
RODEO_REG(chkrebtii_dense) RODEO_REG(chkrebtii_bd) RODEO_REG(fitz_dense) RODEO_REG(fitz_bd)
RODEO_REG(lorenz_dense) RODEO_REG(lorenz_bd) RODEO_REG(mseir_dense) RODEO_REG(mseir_bd)
RODEO_REG(chkrebtii_bds) RODEO_REG(fitz_bds) RODEO_REG(lorenz_bds) RODEO_REG(mseir_bds)
#undef RODEO_REG

void attach_wide(KernelSet *, int);   // kernels_wide.cu

static KernelSet g_sets[64];
static int g_count = -1;
static std::once_flag g_once;

static void fill_sets() {
  int c = 0;
  register_chkrebtii_dense(g_sets, &c);
  register_chkrebtii_bd(g_sets, &c);
  register_fitz_dense(g_sets, &c);
  register_fitz_bd(g_sets, &c);
  register_lorenz_dense(g_sets, &c);
  register_lorenz_bd(g_sets, &c);
  register_mseir_dense(g_sets, &c);
  register_mseir_bd(g_sets, &c);
  register_chkrebtii_bds(g_sets, &c);
  register_fitz_bds(g_sets, &c);
  register_lorenz_bds(g_sets, &c);
  register_mseir_bds(g_sets, &c);
  attach_wide(g_sets, c);
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

const KernelSet *find_kernel_set(int ode, int n, bool blockdiag, bool structured) {
  init_sets();
  for (int i = 0; i < g_count; i++)
    if (g_sets[i].ode == ode && g_sets[i].n == n && g_sets[i].blockdiag == blockdiag &&
        g_sets[i].structured == structured)
      return &g_sets[i];
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
