// kernels_lorenz_bds_p4.cu -- 'lorenz' right-hand side, block-diagonal family specialised for structured priors (upper-triangular Q blocks, w = e_order), block size 4.
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz_bds_p4(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 4, true, true, true>("lorenz/bds-p4");
}
}  // namespace rodeo
