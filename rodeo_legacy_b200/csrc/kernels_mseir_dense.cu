// kernels_mseir_dense.cu -- 'mseir' right-hand side, dense family instantiations.
#include "kernel_launch.cuh"
namespace rodeo {
void register_mseir_dense(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeMseir, 10, false, false>("mseir/n10");
  out[(*count)++] = make_kernel_set<OdeMseir, 15, false, false>("mseir/n15");
}
}  // namespace rodeo
