// kernels_chkrebtii_bds_p3.cu -- 'chkrebtii' right-hand side, block-diagonal family specialised for structured priors (upper-triangular Q blocks, w = e_order), block size 3.
#include "kernel_launch.cuh"
namespace rodeo {
void register_chkrebtii_bds_p3(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 3, true, true, true>("chkrebtii/bds-p3");
}
}  // namespace rodeo
