// kernels_mixed.cu -- dense kernel sets for priors with UNEQUAL block sizes (n_deriv_prior = [3, 4]; the reference
// accepts any list: rodeo/utils/utils.py:97-105, rodeo/ibm/ibm_init.py:54-58).  The right-hand side reads variable
// k at the offset its block starts at (OdeAt, ode_functors.cuh); everything else is the dense family.
// Adding a shape is one line here (plus one attach<> line in kernels_wide.cu when n_state > 6).
#include "kernel_launch.cuh"
namespace rodeo {
void register_mixed(KernelSet *out, int *count) {
  out[(*count)++] = make_kernel_set<OdeAt<OdeFitz, 2, 3>, 5, false, true>("fitz/n5[2,3]");
  out[(*count)++] = make_kernel_set<OdeAt<OdeFitz, 3, 4>, 7, false, false>("fitz/n7[3,4]");
  out[(*count)++] = make_kernel_set<OdeAt<OdeFitz, 4, 3>, 7, false, false>("fitz/n7[4,3]");
  out[(*count)++] = make_kernel_set<OdeAt<OdeLorenz63, 2, 3, 3>, 8, false, false>("lorenz/n8[2,3,3]");
}
}  // namespace rodeo
