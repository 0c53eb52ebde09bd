// kernels_fitz.cu -- kernel instantiations for the 'fitz' right-hand side.
#include "kernel_launch.cuh"
namespace rodeo {
void register_fitz(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeFitz, 4, true>("fitz/n4");
  out[(*count)++] = make_kernel_set<OdeFitz, 6, true>("fitz/n6");
  out[(*count)++] = make_kernel_set<OdeFitz, 8, false>("fitz/n8");
}
}  // namespace rodeo
