// kernels_fitz_bds_p4.cu -- 'fitz' right-hand side, block-diagonal family specialised for structured priors (upper-triangular Q blocks, w = e_order), block size 4.
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz_bds_p4(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 4, true, true, true>("fitz/bds-p4");
}
}  // namespace rodeo
