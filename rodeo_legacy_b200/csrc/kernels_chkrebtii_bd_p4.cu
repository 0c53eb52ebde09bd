// kernels_chkrebtii_bd_p4.cu -- 'chkrebtii' right-hand side, block-diagonal family, block size 4.
#include "kernel_launch.cuh"
namespace rodeo {
void register_chkrebtii_bd_p4(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeChkrebtii, 4, true, true>("chkrebtii/bd-p4");
}
}  // namespace rodeo
