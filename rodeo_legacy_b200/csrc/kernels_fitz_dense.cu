// kernels_fitz_dense.cu -- 'fitz' right-hand side, dense family instantiations.
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz_dense(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 4, false, true>("fitz/n4");
  out[(*count)++] = make_kernel_set<OdeFitz, 6, false, true>("fitz/n6");
  out[(*count)++] = make_kernel_set<OdeFitz, 8, false, false>("fitz/n8");
}
}  // namespace rodeo
