// kernels_chkrebtii_bd_p3.cu -- 'chkrebtii' right-hand side, block-diagonal family, block size 3.
#include "kernel_launch.cuh"
namespace rodeo {
void register_chkrebtii_bd_p3(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 3, true, true>("chkrebtii/bd-p3");
}
}  // namespace rodeo
