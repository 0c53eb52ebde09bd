// ode_functors.cuh -- the registered right-hand sides ("ode_fun" of the reference),
// compiled as device functors instead of a Python callback per step
// (the callback: rodeo/cython/KalmanODE.pyx:323-324).
//
// Formulas: tests/eigen/ode_functions.pyx:8-13 (chkrebtii), :33-42 (FitzHugh-Nagumo,
// written V*V*V/3 as examples/fitz.py:14), :18-28 (Lorenz63), :47-60 (MSEIR); the k-th
// variable's value is X[k*p] with p = n_state / n_var, as in the reference.
//
// Every functor is written against a LAYOUT L (L::off(k) = state index of variable k's value), so that priors
// with unequal block sizes (n_deriv_prior = [3, 4]: rodeo/utils/utils.py:97-105 accepts any list) can be served:
// eval<n> is the reference's equal split, OdeAt<Ode, sizes...> pins the offsets of a block-size pack.
#pragma once
#include "kalman_common.cuh"

namespace rodeo {

// eval<n>: the reference's equal split (variable k at X[k * n / n_var])
#define RODEO_ODE_EQUAL_SPLIT                                                                \
  template <int n>                                                                           \
  RD_HD static void eval(const double *X, double t, const double *th, double *out) {         \
    eval_at<EqualLay<n, n_meas>>(X, t, th, out);                                             \
  }                                                                                          \
  template <int n>                                                                           \
  static constexpr int var_off(int k) { return EqualLay<n, n_meas>::off(k); }                \
  static constexpr bool mixed = false;

template <int n, int m>
struct EqualLay {
  RD_HD static constexpr int off(int k) { return k * (n / m); }
};
template <int... S>
struct BlockLay {
  static constexpr int m = sizeof...(S);
  RD_HD static constexpr int off(int k) {
    constexpr int sz[m] = {S...};
    int o = 0;
    for (int i = 0; i < k; i++) o += sz[i];
    return o;
  }
  static constexpr int n = (S + ...);
};

struct OdeChkrebtii {
  static constexpr int id = 0, n_meas = 1, n_theta = 0, order = 2;   // x'' = ...: W picks the 2nd derivative
  template <class L>
  RD_HD static void eval_at(const double *X, double t, const double *, double *out) {
    out[0] = sin(2 * t) - X[0];
  }
  RODEO_ODE_EQUAL_SPLIT
};

struct OdeFitz {
  static constexpr int id = 1, n_meas = 2, n_theta = 3, order = 1;
  template <class L>
  RD_HD static void eval_at(const double *X, double t, const double *th, double *out) {
    const double a = th[0], b = th[1], c = th[2];
    constexpr int o0 = L::off(0), o1 = L::off(1);
    const double V = X[o0], R = X[o1];
    out[0] = c * (V - V * V * V / 3 + R);
    out[1] = -1 / c * (V - a + b * R);
  }
  RODEO_ODE_EQUAL_SPLIT
};

struct OdeLorenz63 {
  static constexpr int id = 2, n_meas = 3, n_theta = 3, order = 1;
  template <class L>
  RD_HD static void eval_at(const double *X, double t, const double *th, double *out) {
    const double rho = th[0], sigma = th[1], beta = th[2];
    constexpr int o0 = L::off(0), o1 = L::off(1), o2 = L::off(2);
    const double x = X[o0], y = X[o1], z = X[o2];
    out[0] = -sigma * x + sigma * y;
    out[1] = rho * x - y - x * z;
    out[2] = -beta * z + x * y;
  }
  RODEO_ODE_EQUAL_SPLIT
};

struct OdeMseir {
  static constexpr int id = 3, n_meas = 5, n_theta = 6, order = 1;
  template <class L>
  RD_HD static void eval_at(const double *X, double t, const double *th, double *out) {
    constexpr int o0 = L::off(0), o1 = L::off(1), o2 = L::off(2), o3 = L::off(3), o4 = L::off(4);
    const double M = X[o0], S = X[o1], E = X[o2], I = X[o3], R = X[o4];
    const double N = M + S + E + I + R;
    const double Lambda = th[0], delta = th[1], beta = th[2], mu = th[3], epsilon = th[4],
                 gamma = th[5];
    out[0] = Lambda - delta * M - mu * M;
    out[1] = delta * M - beta * S * I / N - mu * S;
    out[2] = beta * S * I / N - (epsilon + mu) * E;
    out[3] = epsilon * E - (gamma + mu) * I;
    out[4] = gamma * I - mu * R;
  }
  RODEO_ODE_EQUAL_SPLIT
};

// The same right-hand side with the variables at the offsets of the block sizes S... (sum = n_state).
template <class Ode, int... S>
struct OdeAt : Ode {
  using Lay = BlockLay<S...>;
  static_assert(Lay::m == Ode::n_meas, "one block size per ODE variable");
  template <int n>
  RD_HD static void eval(const double *X, double t, const double *th, double *out) {
    static_assert(n == Lay::n, "block sizes must add up to n_state");
    Ode::template eval_at<Lay>(X, t, th, out);
  }
  template <int n>
  static constexpr int var_off(int k) { return Lay::off(k); }
  static constexpr bool mixed = true;
};
#undef RODEO_ODE_EQUAL_SPLIT

}  // namespace rodeo
