#!/usr/bin/env python
"""Stress: repeat full-size solves and require bit-identical results.  With a shared prior the covariance path
does not depend on theta / x0, so var must be bit-identical across the whole batch and across repetitions; mu
must repeat bit-for-bit.  (This property test is what exposed the cross-proxy ring race)

    python tools/stress_bit_identity.py [--reps 100]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rodeo_legacy_b200 import workloads as wl


def check_sim(name, maker, B, reps):
    """solve_sim (the draw-only thread-per-block kernel for >= 3 blocks): draws must repeat bit for bit."""
    w, x0, th = maker(B)
    n = len(w["x0"])
    z = np.asfortranarray(np.random.default_rng(1).standard_normal((n, w["N"])))
    k = wl.make_kode(w, z_state=z)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(th).cuda()
    ref, bad = None, 0
    for r in range(reps):
        x = k.solve_sim(x0d, theta=thd)
        if ref is None:
            ref = x.clone()
        bad += not bool(torch.equal(x, ref))
        del x
    print(f"{name}: {reps} repetitions, {bad} mismatching", flush=True)
    return bad


def check(name, maker, B, reps, mode="mv", **kw):
    w, x0, th = maker(B)
    n = len(w["x0"])
    z = np.asfortranarray(np.random.default_rng(1).standard_normal((n, w["N"])))
    k = wl.make_kode(w, z_state=z, **kw)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(th).cuda()
    ref_mu = ref_var0 = None
    bad = 0
    for r in range(reps):
        if mode == "mv":
            mu, var = k.solve_mv(x0d, theta=thd)
        else:
            _, mu, var = k.solve(x0d, theta=thd)
        same = True
        for lo in range(0, B, 4096):        # var identical across the batch (chunked compare: memory)
            same &= bool((var[lo:lo + 4096] == var[0:1]).all())
        if ref_mu is None:
            ref_mu, ref_var0 = mu.clone(), var[0].clone()
        same &= bool(torch.equal(mu, ref_mu)) and bool(torch.equal(var[0], ref_var0))
        bad += (not same)
        del mu, var
    print(f"{name}: {reps} repetitions, {bad} mismatching", flush=True)
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=100)
    a = ap.parse_args()
    bad = check("fitz 65536 x 400 solve_mv", wl.fitz_batch, 65536, a.reps)
    bad += check("fitz 65536 x 400 solve", wl.fitz_batch, 65536, max(1, a.reps // 4), mode="both")
    bad += check("mseir 8192 x 1000 solve_mv", lambda B: wl.mseir_batch(B, N=1000), 8192, max(1, a.reps // 4))
    bad += check("lorenz 8192 x 1000 solve", lambda B: wl.lorenz_batch(B, N=1000, tmax=4.0), 8192, max(1, a.reps // 4), mode="both")
    bad += check("fitz 65536 x 400 solve_mv, var_layout blocks", wl.fitz_batch, 65536, max(1, a.reps // 4), var_layout="blocks")
    bad += check("fitz 65519 x 401 solve_mv (ragged batch, odd grid)",
                 lambda B: (lambda t: (dict(t[0], N=401, tmax=40.1), t[1], t[2]))(wl.fitz_batch(B)), 65519, max(1, a.reps // 4))
    bad += check_sim("lorenz 16384 x 5000 solve_sim", wl.lorenz_batch, 16384, max(1, a.reps // 4))
    sys.exit(1 if bad else 0)
