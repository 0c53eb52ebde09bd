// kernels_lorenz_dense.cu -- 'lorenz' right-hand side, dense family instantiations.
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz_dense(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 6, false, true>("lorenz/n6");
  out[(*count)++] = make_kernel_set<OdeLorenz63, 9, false, false>("lorenz/n9");
}
}  // namespace rodeo
