// kernels_lorenz_bd_p3.cu -- 'lorenz' right-hand side, block-diagonal family, block size 3.
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz_bd_p3(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 3, true, true>("lorenz/bd-p3");
}
}  // namespace rodeo
