// ode_functors.cuh -- the registered right-hand sides ("ode_fun" of the reference),
// compiled as device functors instead of a Python callback per step
// (the callback: rodeo/cython/KalmanODE.pyx:323-324).
//
// Formulas: tests/eigen/ode_functions.pyx:8-13 (chkrebtii), :33-42 (FitzHugh-Nagumo,
// written V*V*V/3 as examples/fitz.py:14), :18-28 (Lorenz63), :47-60 (MSEIR); the k-th
// variable's value is X[k*p] with p = n_state / n_var, as in the reference.
#pragma once
#include "kalman_common.cuh"

namespace rodeo {

struct OdeChkrebtii {
  static constexpr int id = 0, n_meas = 1, n_theta = 0, order = 2;   // x'' = ...: W picks the 2nd derivative
  template <int n>
  RD_HD static void eval(const double *X, double t, const double *, double *out) {
    out[0] = sin(2 * t) - X[0];
  }
};

struct OdeFitz {
  static constexpr int id = 1, n_meas = 2, n_theta = 3, order = 1;
  template <int n>
  RD_HD static void eval(const double *X, double t, const double *th, double *out) {
    constexpr int p = n / 2;
    const double a = th[0], b = th[1], c = th[2];
    const double V = X[0], R = X[p];
    out[0] = c * (V - V * V * V / 3 + R);
    out[1] = -1 / c * (V - a + b * R);
  }
};

struct OdeLorenz63 {
  static constexpr int id = 2, n_meas = 3, n_theta = 3, order = 1;
  template <int n>
  RD_HD static void eval(const double *X, double t, const double *th, double *out) {
    constexpr int p = n / 3;
    const double rho = th[0], sigma = th[1], beta = th[2];
    const double x = X[0], y = X[p], z = X[2 * p];
    out[0] = -sigma * x + sigma * y;
    out[1] = rho * x - y - x * z;
    out[2] = -beta * z + x * y;
  }
};

struct OdeMseir {
  static constexpr int id = 3, n_meas = 5, n_theta = 6, order = 1;
  template <int n>
  RD_HD static void eval(const double *X, double t, const double *th, double *out) {
    constexpr int p = n / 5;
    const double M = X[0], S = X[p], E = X[2 * p], I = X[3 * p], R = X[4 * p];
    const double N = M + S + E + I + R;
    const double Lambda = th[0], delta = th[1], beta = th[2], mu = th[3], epsilon = th[4],
                 gamma = th[5];
    out[0] = Lambda - delta * M - mu * M;
    out[1] = delta * M - beta * S * I / N - mu * S;
    out[2] = beta * S * I / N - (epsilon + mu) * E;
    out[3] = epsilon * E - (gamma + mu) * I;
    out[4] = gamma * I - mu * R;
  }
};

}  // namespace rodeo
