#!/usr/bin/env python
"""Time every BASELINE.json configuration on one GPU (device-resident inputs/outputs).
Not the headline bench (that is bench.py); a table for DESIGN.md / profiles.

    python tools/bench_configs.py [--scale 1.0] [--reps 3]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from rodeo_legacy_b200 import workloads as wl


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--checkpoint", type=int, default=0)
    ap.add_argument("--dense", action="store_true", help="force the general dense kernel families")
    a = ap.parse_args()
    rows = []

    def run(name, maker, B, mode, n_alg_bytes_step):
        B = max(32, int(B * a.scale))
        w, x0, theta = maker(B)
        n = len(w["x0"])
        z = None
        if mode not in ("mv", "mv_pp"):
            np.random.seed(2)
            z = np.asfortranarray(np.random.randn(n, w["N"]))
        k = wl.make_kode(w, z_state=z, checkpoint=a.checkpoint, force_dense=a.dense)
        x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
        k.set_timing(True)
        if mode == "mv":
            fn = lambda: k.solve_mv(x0d, theta=thd)
        elif mode == "sim":
            fn = lambda: k.solve_sim(x0d, theta=thd)
        elif mode == "both":
            fn = lambda: k.solve(x0d, theta=thd)
        elif mode == "mv_pp":       # every system its own var_state (a sweep over the prior scale sigma)
            scale = torch.linspace(0.5, 2.0, B, dtype=torch.float64, device="cuda")
            Rb = torch.from_numpy(np.ascontiguousarray(w["R"])).cuda()[None] * scale[:, None, None] ** 2
            fn = lambda: k.solve_mv(x0d, theta=thd, priors=dict(var_state=Rb))
        else:
            ind = np.arange(w["N"] // 40, w["N"] + 1, w["N"] // 40)
            Y = np.zeros((len(ind), 2))
            fn = lambda: k.loglik(x0d, thd, Y, ind, [0, 3], 0.2)
        ms = timed(fn, a.reps)
        f, b = k.last_kernel_ms()
        rows.append(dict(config=name, family="/".join(k.kernel_family), checkpoint=k.checkpoint, batch=B, n_state=n, n_eval=w["N"], mode=mode, ms=ms,
                         solves_per_s=B / ms * 1e3, forward_ms=f, backward_ms=b,
                         alg_GBps=n_alg_bytes_step * w["N"] * B / ms / 1e6,
                         launches=k.launch_count))
        print(json.dumps(rows[-1]), flush=True)
        del k
        torch.cuda.empty_cache()

    hist = lambda n: 8 * 2 * (2 * n + 2 * n * n)
    sel = a.only.split(",") if a.only else None
    cfgs = [
        ("C2 fitz solve_mv", wl.fitz_batch, 65536, "mv", hist(6) + 8 * (6 + 36)),
        ("C3 lorenz63 solve_sim", wl.lorenz_batch, 16384, "sim", hist(9) + 8 * 9),
        ("C4 mseir solve_mv", wl.mseir_batch, 32768, "mv", hist(15) + 8 * (15 + 225)),
        ("C5 fitz loglik (per GPU of 8)", wl.fitz_batch, 131072, "ll", hist(6)),
        ("C2pp fitz solve_mv, per-system var_state", wl.fitz_batch, 65536, "mv_pp", hist(6) + 8 * (6 + 36)),
        ("C2sim fitz solve_sim", wl.fitz_batch, 65536, "sim", hist(6) + 8 * 6),
        ("C2both fitz solve", wl.fitz_batch, 65536, "both", hist(6) + 8 * (12 + 36)),
        ("C4sim mseir solve_sim", wl.mseir_batch, 32768, "sim", hist(15) + 8 * 15),
    ]
    for c in cfgs:
        if sel is None or any(s in c[0] for s in sel):
            run(*c)


if __name__ == "__main__":
    main()
