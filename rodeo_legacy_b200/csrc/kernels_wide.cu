// kernels_wide.cu -- dense wide-state kernels (half a warp per system, kalman_dw.cuh) attached to
// the dense kernel sets whose state does not fit one thread's registers.
#include "kernel_launch.cuh"
#include "kalman_dw.cuh"
namespace rodeo {

template <class Ode, int n>
static void attach(KernelSet *sets, int count) {
  using Fam = DwFam<Ode, n>;
  for (int i = 0; i < count; i++) {
    KernelSet &k = sets[i];
    if (k.ode != Ode::id || k.n != n || k.blockdiag || k.mixed != Ode::mixed) continue;
    bool same = true;
    for (int a = 0; a < Ode::n_meas; a++) same = same && k.off[a] == Ode::template var_off<n>(a);
    if (!same) continue;
    k.wide_forward[IG_PROBDE] = launch_dw_forward<Fam, IG_PROBDE>;
    k.wide_forward[IG_SCHOBERT] = launch_dw_forward<Fam, IG_SCHOBERT>;
    k.wide_forward[IG_CHKREBTII] = launch_dw_forward<Fam, IG_CHKREBTII>;
    k.wide_backward[BK_MV] = launch_dw_backward<Fam, true, true, false, false>;
    k.wide_backward[BK_SIM] = launch_dw_backward<Fam, false, false, true, false>;
    k.wide_backward[BK_BOTH] = launch_dw_backward<Fam, true, true, true, false>;
    k.wide_backward[BK_LL_MEAN] = launch_dw_backward<Fam, true, false, false, true>;
    k.wide_backward[BK_LL_DRAW] = launch_dw_backward<Fam, false, false, true, true>;
    k.wide_pp_forward = launch_dw_forward_pp<Fam>;
    k.wide_pp_backward[BK_MV] = launch_dw_backward_pp<Fam, true, true, false, false>;
    k.wide_pp_backward[BK_SIM] = launch_dw_backward_pp<Fam, false, false, true, false>;
    k.wide_pp_backward[BK_BOTH] = launch_dw_backward_pp<Fam, true, true, true, false>;
    k.wide_pp_backward[BK_LL_MEAN] = launch_dw_backward_pp<Fam, true, false, false, true>;
    k.wide_pp_backward[BK_LL_DRAW] = launch_dw_backward_pp<Fam, false, false, true, true>;
    k.wide_hist_elems = Fam::NE;
  }
}

void attach_wide(KernelSet *sets, int count) {
  attach<OdeFitz, 8>(sets, count);
  attach<OdeLorenz63, 9>(sets, count);
  attach<OdeMseir, 10>(sets, count);
  attach<OdeMseir, 15>(sets, count);
  attach<OdeAt<OdeFitz, 3, 4>, 7>(sets, count);       // unequal block sizes (kernels_mixed.cu)
  attach<OdeAt<OdeFitz, 4, 3>, 7>(sets, count);
  attach<OdeAt<OdeLorenz63, 2, 3, 3>, 8>(sets, count);
}

}  // namespace rodeo
