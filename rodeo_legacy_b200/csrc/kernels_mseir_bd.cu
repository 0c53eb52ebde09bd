// kernels_mseir_bd.cu -- 'mseir' right-hand side, bd family instantiations.
#include "kernel_launch.cuh"
namespace rodeo {
void register_mseir_bd(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeMseir, 2, true, true>("mseir/bd-p2");
  out[(*count)++] = make_kernel_set<OdeMseir, 3, true, true>("mseir/bd-p3");
}
}  // namespace rodeo
