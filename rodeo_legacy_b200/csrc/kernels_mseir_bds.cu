// kernels_mseir_bds.cu -- 'mseir' right-hand side, block-diagonal family specialised for structured priors
// (upper-triangular Q blocks, w = e_order: what ibm_init + zero_pad produce; kalman_math.inl).
#include "kernel_launch.cuh"
namespace rodeo {
void register_mseir_bds(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeMseir, 2, true, true, true>("mseir/bds-p2");
  out[(*count)++] = make_kernel_set<OdeMseir, 3, true, true, true>("mseir/bds-p3");
}
}  // namespace rodeo
