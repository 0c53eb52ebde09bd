// kernels_lorenz_bd.cu -- 'lorenz' right-hand side, bd family instantiations.
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz_bd(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 2, true, true>("lorenz/bd-p2");
  out[(*count)++] = make_kernel_set<OdeLorenz63, 3, true, true>("lorenz/bd-p3");
  out[(*count)++] = make_kernel_set<OdeLorenz63, 4, true, true>("lorenz/bd-p4");
}
}  // namespace rodeo
