// registry.cu -- collects the compiled kernel sets.
#include "registry.cuh"

namespace rodeo {

void register_chkrebtii(KernelSet *, int *);
void register_fitz(KernelSet *, int *);
void register_lorenz(KernelSet *, int *);
void register_mseir(KernelSet *, int *);

static KernelSet g_sets[32];
static int g_count = -1;

static void init_sets() {
  if (g_count >= 0) return;
  int c = 0;
  register_chkrebtii(g_sets, &c);
  register_fitz(g_sets, &c);
  register_lorenz(g_sets, &c);
  register_mseir(g_sets, &c);
  g_count = c;
}

const KernelSet *find_kernel_set(int ode, int n) {
  init_sets();
  for (int i = 0; i < g_count; i++)
    if (g_sets[i].ode == ode && g_sets[i].n == n) return &g_sets[i];
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
