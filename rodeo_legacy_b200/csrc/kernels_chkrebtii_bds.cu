// kernels_chkrebtii_bds.cu -- 'chkrebtii' right-hand side, block-diagonal family specialised for structured priors
// (upper-triangular Q blocks, w = e_order: what ibm_init + zero_pad produce; kalman_math.inl).
#include "kernel_launch.cuh"
namespace rodeo {
void register_chkrebtii_bds(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 3, true, true, true>("chkrebtii/bds-p3");
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 4, true, true, true>("chkrebtii/bds-p4");
}
}  // namespace rodeo
