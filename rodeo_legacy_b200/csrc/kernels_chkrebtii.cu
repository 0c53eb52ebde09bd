// kernels_chkrebtii.cu -- kernel instantiations for the 'chkrebtii' right-hand side.
#include "kernel_launch.cuh"
namespace rodeo {
void register_chkrebtii(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 3, true>("chkrebtii/n3");
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 4, true>("chkrebtii/n4");
}
}  // namespace rodeo
