// kernels_lorenz_bds.cu -- 'lorenz' right-hand side, block-diagonal family specialised for structured priors
// (upper-triangular Q blocks, w = e_order: what ibm_init + zero_pad produce; kalman_math.inl).
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz_bds(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 2, true, true, true>("lorenz/bds-p2");
  out[(*count)++] = make_kernel_set<OdeLorenz63, 3, true, true, true>("lorenz/bds-p3");
  out[(*count)++] = make_kernel_set<OdeLorenz63, 4, true, true, true>("lorenz/bds-p4");
}
}  // namespace rodeo
