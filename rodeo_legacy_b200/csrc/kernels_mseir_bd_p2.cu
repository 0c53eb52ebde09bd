// kernels_mseir_bd_p2.cu -- 'mseir' right-hand side, block-diagonal family, block size 2.
#include "kernel_launch.cuh"
namespace rodeo {
void register_mseir_bd_p2(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeMseir, 2, true, true>("mseir/bd-p2");
}
}  // namespace rodeo
