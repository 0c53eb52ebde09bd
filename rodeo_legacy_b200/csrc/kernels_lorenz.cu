// kernels_lorenz.cu -- kernel instantiations for the 'lorenz' right-hand side.
#include "kernel_launch.cuh"
namespace rodeo {
void register_lorenz(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeLorenz63, 6, true>("lorenz/n6");
  out[(*count)++] = make_kernel_set<OdeLorenz63, 9, false>("lorenz/n9");
}
}  // namespace rodeo
