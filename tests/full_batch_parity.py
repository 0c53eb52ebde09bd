#!/usr/bin/env python
"""One-off (not collected by pytest): EVERY system of the BASELINE configurations against the CPU oracle.

The `-m gpu` tests compare samples of the full-size batches (the oracle must finish in seconds); this script
takes the minutes needed to compare all of them, chunk by chunk, and prints the worst trajectory-scaled error.
Lives under tests/ because it uses the oracle (test infrastructure).

    python tests/full_batch_parity.py [--configs C2,C3,C4] [--chunk 1024] > profiles/rNN_full_batch_parity.jsonl
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from tests import workloads as wl
from tests.common import traj_err, var_err


def run(name, maker, B, mode, chunk, horizon=None, dense=False):
    w, x0, th = maker(B)
    n = len(w["x0"])
    np.random.seed(2)
    z = np.asfortranarray(np.random.randn(n, w["N"])) if mode == "sim" else None
    k = wl.make_kode(w, z_state=z, force_dense=dense)
    po = wl.make_oracle_problem(w)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(th).cuda()
    if mode == "mv":
        mu, var = k.solve_mv(x0d, theta=thd)
    else:
        xs = k.solve_sim(x0d, theta=thd)
    status_bad = int(k.status.abs().sum())
    worst = {}
    t0 = time.perf_counter()
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        ref = oracle.solve_batch(po, x0[lo:hi], th[lo:hi], z, oracle.MODE_MV if mode == "mv" else oracle.MODE_SIM,
                                 len(os.sched_getaffinity(0)))
        if mode == "mv":
            e = {"mu": traj_err(mu[lo:hi].cpu().numpy(), ref["mu"]).max(),
                 "var": var_err(var[lo:hi].cpu().numpy()[:, 1:], ref["var"][:, 1:]).max()}
        else:
            g = xs[lo:hi].cpu().numpy()
            e = {"x": traj_err(g, ref["x"]).max()}
            if horizon:
                e["x_first_%d_steps" % horizon] = traj_err(g[:, :horizon], ref["x"][:, :horizon]).max()
        for kk, v in e.items():
            worst[kk] = max(worst.get(kk, 0.0), float(v))
    print(json.dumps({"config": name, "family": "/".join(k.kernel_family), "systems_compared": B, "n_eval": w["N"],
                      "mode": mode, "worst_trajectory_scaled_error": worst, "status_nonzero": status_bad,
                      "oracle_seconds": round(time.perf_counter() - t0, 1)}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C3,C4")
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--dense", action="store_true", help="force the general dense kernel families")
    a = ap.parse_args()
    sel = a.configs.split(",")
    if "C2" in sel:
        run("C2 fitz solve_mv 65536 x 400", wl.fitz_batch, 65536, "mv", a.chunk, dense=a.dense)
    if "C4" in sel:
        run("C4 mseir solve_mv 32768 x 1000", wl.mseir_batch, 32768, "mv", min(a.chunk, 512), dense=a.dense)
    if "C3" in sel:
        run("C3 lorenz63 solve_sim 16384 x 5000", wl.lorenz_batch, 16384, "sim", a.chunk, horizon=1000, dense=a.dense)
