// kernels_fitz_bd_p4.cu -- 'fitz' right-hand side, block-diagonal family, block size 4.
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz_bd_p4(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 4, true, true>("fitz/bd-p4");
}
}  // namespace rodeo
