// kernels_chkrebtii_dense.cu -- 'chkrebtii' right-hand side, dense family instantiations.
#include "kernel_launch.cuh"
namespace rodeo {
void register_chkrebtii_dense(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 3, false, true>("chkrebtii/n3");
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 4, false, true>("chkrebtii/n4");
}
}  // namespace rodeo
