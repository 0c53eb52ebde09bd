// kernels_fitz_bd_p2.cu -- 'fitz' right-hand side, block-diagonal family, block size 2.
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz_bd_p2(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 2, true, true>("fitz/bd-p2");
}
}  // namespace rodeo
