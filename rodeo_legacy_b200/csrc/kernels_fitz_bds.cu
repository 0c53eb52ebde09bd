// kernels_fitz_bds.cu -- 'fitz' right-hand side, block-diagonal family specialised for structured priors
// (upper-triangular Q blocks, w = e_order: what ibm_init + zero_pad produce; kalman_math.inl).
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz_bds(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 2, true, true, true>("fitz/bds-p2");
  out[(*count)++] = make_kernel_set<OdeFitz, 3, true, true, true>("fitz/bds-p3");
  out[(*count)++] = make_kernel_set<OdeFitz, 4, true, true, true>("fitz/bds-p4");
}
}  // namespace rodeo
