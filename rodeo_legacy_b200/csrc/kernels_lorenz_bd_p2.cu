// kernels_lorenz_bd_p2.cu -- 'lorenz' right-hand side, block-diagonal family, block size 2.
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz_bd_p2(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 2, true, true>("lorenz/bd-p2");
}
}  // namespace rodeo
