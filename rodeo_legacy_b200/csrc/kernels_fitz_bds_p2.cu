// kernels_fitz_bds_p2.cu -- 'fitz' right-hand side, block-diagonal family specialised for structured priors (upper-triangular Q blocks, w = e_order), block size 2.
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz_bds_p2(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 2, true, true, true>("fitz/bds-p2");
}
}  // namespace rodeo
