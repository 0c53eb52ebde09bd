#!/usr/bin/env python
"""FitzHugh-Nagumo parameter inference, the batched way.

The reference (examples/fitz.py + examples/inference.py) finds the posterior mode of
phi = log(theta) with Nelder-Mead and a numerical Hessian: thousands of SEQUENTIAL calls of
`kalman_nlpost`, each one `KalmanODE.solve_sim` -> `thinning` -> Gaussian log-likelihood
(examples/inference.py:85-94).  Here the same negative log-posterior is evaluated for a whole
cloud of phi at once -- one forward + one backward kernel, 8 bytes per theta coming back -- and the
Laplace approximation (mode + covariance of phi) is read off the cloud by importance weighting.

    python examples/fitz_likelihood_sweep.py [--n 262144]

Needs a CUDA device (there is no CPU fallback).  Uses only the public API:
    ibm_init / indep_init / zero_pad  and  rodeo_legacy_b200.cuda.KalmanODE(...).loglik(...)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from rodeo_legacy_b200 import ibm_init, indep_init, zero_pad
from rodeo_legacy_b200.cuda import KalmanODE


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 18, help="number of theta draws in the sweep")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()

    # ---- problem definition: examples/fitz.py:18-53 ------------------------------------------
    n_deriv, n_deriv_prior = [1, 1], [3, 3]
    state_ind = [0, 3]
    tmin, tmax, h = 0.0, 40.0, 0.1
    n_eval = int((tmax - tmin) / h)
    W = zero_pad(np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]), n_deriv, n_deriv_prior)
    x0_state = zero_pad(np.array([-1.0, 1.0, 1.0, 1 / 3]), n_deriv, n_deriv_prior)
    theta_true = np.array([0.2, 0.2, 3.0])
    phi_mean, phi_sd, gamma = np.log(theta_true), np.ones(3), 0.2
    kinit = indep_init(ibm_init(h, n_deriv_prior, [0.1, 0.1]), n_deriv_prior)
    kode = KalmanODE(W, tmin, tmax, n_eval, "fitz", **kinit)

    # ---- synthetic observations at t = 1..40 (inference.simulate: odeint + noise; here the
    # solver's own posterior mean at theta_true stands in for odeint) --------------------------
    rng = np.random.default_rng(a.seed)
    mu_true, _ = kode.solve_mv(x0_state, theta=theta_true)
    thin_index = np.arange(1, int(tmax) + 1) * int(round(1 / h))         # grid index of each data time
    Y = mu_true[thin_index][:, state_ind] + gamma * rng.standard_normal((len(thin_index), 2))

    # ---- one sweep: log-posterior of a cloud of phi --------------------------------------------
    phi = phi_mean + 0.15 * rng.standard_normal((a.n, 3))
    theta = torch.from_numpy(np.exp(phi)).cuda()
    X0 = torch.from_numpy(np.tile(x0_state, (a.n, 1))).cuda()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kode.loglik(X0[:1024], theta[:1024], Y, thin_index, state_ind, gamma)   # warm-up
    t0.record()
    ll = kode.loglik(X0, theta, Y, thin_index, state_ind, gamma)            # (n,) on the GPU
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    logprior = -0.5 * np.sum(((phi - phi_mean) / phi_sd) ** 2, axis=1)
    lp = ll.cpu().numpy() + logprior

    # ---- Laplace-style summary from the cloud (proposal N(phi_mean, 0.15^2 I)) ------------------
    logq = -0.5 * np.sum(((phi - phi_mean) / 0.15) ** 2, axis=1)
    w = np.exp((lp - logq) - np.max(lp - logq))
    w /= w.sum()
    mode = phi[np.argmax(lp)]
    mean = w @ phi
    cov = (phi - mean).T @ ((phi - mean) * w[:, None])
    print(f"{a.n} solves + likelihoods in {ms:.2f} ms ({a.n / ms * 1e3 / 1e6:.1f} M solves/s), "
          f"checkpoint {kode.checkpoint}, family {kode.kernel_family}")
    print("theta true          ", theta_true)
    print("theta at the mode   ", np.exp(mode))
    print("posterior mean theta", np.exp(mean))
    print("posterior sd of phi ", np.sqrt(np.diag(cov)))
    print("effective sample size", 1.0 / np.sum(w ** 2))


if __name__ == "__main__":
    main()
