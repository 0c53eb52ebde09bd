// kernels_fitz_bd.cu -- 'fitz' right-hand side, bd family instantiations.
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz_bd(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 2, true, true>("fitz/bd-p2");
  out[(*count)++] = make_kernel_set<OdeFitz, 3, true, true>("fitz/bd-p3");
  out[(*count)++] = make_kernel_set<OdeFitz, 4, true, true>("fitz/bd-p4");
}
}  // namespace rodeo
