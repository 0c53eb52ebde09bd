#!/usr/bin/env python
"""Tuning the prior scale sigma of the probabilistic ODE solver, the batched way.

The IBM prior of rodeo has one free scale per variable (`ibm_init(dt, n_deriv_prior, sigma)`); the reference's
examples fix it by hand (examples/fitz.py: sigma = [.1, .1]).  Choosing it from data means solving the SAME ODE
under MANY priors -- in the reference one `KalmanODE` object and one sequential solve per candidate.  Here every
candidate sigma is one system of a single batched call: `KalmanODE.loglik(..., priors=dict(var_state=R_b))` gives the
thinned Gaussian log-likelihood of the solution posterior mean for each prior without materialising a trajectory
(`rodeo_cuda_loglik_batched_prior`), and `solve_mv(..., priors=...)` returns the posterior of the winners
(`rodeo_cuda_solve_batched_prior`: block-diagonal per-system priors run on the thread-per-block kernels).

    python examples/sigma_sweep.py [--n 4096]

Needs a CUDA device (there is no CPU fallback).
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
    ap.add_argument("--n", type=int, default=4096, help="number of candidate (sigma_V, sigma_R) pairs")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()

    # ---- problem definition: examples/fitz.py:18-53 ------------------------------------------
    n_deriv, n_deriv_prior = [1, 1], [3, 3]
    tmin, tmax, n_eval = 0.0, 40.0, 400
    dt = (tmax - tmin) / n_eval
    W = zero_pad(np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0]]), n_deriv, n_deriv_prior)
    x0 = zero_pad(np.array([-1.0, 1.0, 1.0, 1 / 3]), n_deriv, n_deriv_prior)
    theta = np.array([0.2, 0.2, 3.0])
    kinit = indep_init(ibm_init(dt, n_deriv_prior, [0.1, 0.1]), n_deriv_prior)
    kode = KalmanODE(W, tmin, tmax, n_eval, "fitz", **kinit)

    # ---- synthetic observations at t = 1..40 (examples/inference.py:58-64): a fine-grid solve + noise ----------
    rng = np.random.default_rng(a.seed)
    fine = 4000
    kfine = KalmanODE(W, tmin, tmax, fine, "fitz",
                      **indep_init(ibm_init((tmax - tmin) / fine, n_deriv_prior, [0.1, 0.1]), n_deriv_prior))
    mu_fine, _ = kfine.solve_mv(x0, theta=theta)
    obs_t = np.arange(1, 41)
    gamma = 0.2
    Y = mu_fine[obs_t * (fine // 40)][:, [0, 3]] + gamma * rng.standard_normal((40, 2))
    thin = obs_t * (n_eval // 40)

    # ---- the sweep: one prior per system ------------------------------------------------------------------------
    B = a.n
    sig = np.exp(rng.uniform(np.log(1e-3), np.log(10.0), size=(B, 2)))
    R = np.stack([indep_init(ibm_init(dt, n_deriv_prior, list(s)), n_deriv_prior)["var_state"] for s in sig])
    X0 = torch.from_numpy(np.tile(x0, (B, 1))).cuda()
    TH = torch.from_numpy(np.tile(theta, (B, 1))).cuda()
    pri = dict(var_state=torch.from_numpy(R).cuda())
    ll = kode.loglik(X0, TH, Y, thin, [0, 3], gamma, priors=pri)
    best = int(ll.argmax())
    print(f"{B} candidate priors in one call; best sigma = ({sig[best, 0]:.3g}, {sig[best, 1]:.3g}), "
          f"log-likelihood {float(ll[best]):.2f} (hand-picked (0.1, 0.1): "
          f"{float(kode.loglik(X0[:1], TH[:1], Y, thin, [0, 3], gamma)[0]):.2f})")

    # ---- posterior under the ten best priors -----------------------------------------------------------------------
    top = torch.topk(ll, 10).indices
    mu, var = kode.solve_mv(X0[top], theta=TH[top], priors=dict(var_state=pri["var_state"][top]))
    print("ran on the block-diagonal kernels:", kode.last_batched_prior_blockdiag,
          "| posterior sd of V at t = 40 under the ten best priors:",
          np.round(np.sqrt(var[:, -1, 0, 0].cpu().numpy()), 4))


if __name__ == "__main__":
    main()
