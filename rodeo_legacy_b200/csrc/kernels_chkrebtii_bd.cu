// kernels_chkrebtii_bd.cu -- 'chkrebtii' right-hand side, bd family instantiations.
#include "kernel_launch.cuh"
namespace rodeo {
void register_chkrebtii_bd(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 3, true, true>("chkrebtii/bd-p3");
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 4, true, true>("chkrebtii/bd-p4");
}
}  // namespace rodeo
