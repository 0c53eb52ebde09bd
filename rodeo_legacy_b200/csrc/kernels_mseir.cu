// kernels_mseir.cu -- kernel instantiations for the 'mseir' right-hand side.
#include "kernel_launch.cuh"
namespace rodeo {
void register_mseir(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeMseir, 10, false>("mseir/n10");
  out[(*count)++] = make_kernel_set<OdeMseir, 15, false>("mseir/n15");
}
}  // namespace rodeo
