// kernels_fitz_bd_p3.cu -- 'fitz' right-hand side, block-diagonal family, block size 3.
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz_bd_p3(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 3, true, true>("fitz/bd-p3");
}
}  // namespace rodeo
