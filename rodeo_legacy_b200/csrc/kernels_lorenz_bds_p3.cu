// kernels_lorenz_bds_p3.cu -- 'lorenz' right-hand side, block-diagonal family specialised for structured priors (upper-triangular Q blocks, w = e_order), block size 3.
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz_bds_p3(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 3, true, true, true>("lorenz/bds-p3");
}
}  // namespace rodeo
