// kernels_lorenz_bd_p4.cu -- 'lorenz' right-hand side, block-diagonal family, block size 4.
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz_bd_p4(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 4, true, true>("lorenz/bd-p4");
}
}  // namespace rodeo
