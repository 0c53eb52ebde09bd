// kernels_mseir_bd_p3.cu -- 'mseir' right-hand side, block-diagonal family, block size 3.
#include "kernel_launch.cuh"
namespace rodeo {
void register_mseir_bd_p3(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeMseir, 3, true, true>("mseir/bd-p3");
}
}  // namespace rodeo
