#!/usr/bin/env python
"""BASELINE config 5: FitzHugh-Nagumo log-likelihood sweep over 1,048,576 theta draws, sharded
over the GPUs of one box (one rank per GPU under torchrun), no data-path collective, ONE NCCL
all_gather of 8 bytes per system at the end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/sweep_c5.py [--total 1048576]

Prints one JSON line from rank 0 (device time = max over ranks, gather included).  The values are
checked against the CPU oracle in tests/test_gpu_parity.py::test_config5_sweep (this tool never
touches the oracle)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from rodeo_legacy_b200.dist import gather_batch, shard_range
from rodeo_legacy_b200 import workloads as wl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = wl.fitz()
    # the same global theta stream on every rank; each rank takes its contiguous slice
    lo, hi = shard_range(a.total, rank, world)
    rng = np.random.default_rng(0)
    eps = rng.standard_normal((a.total, 3))[lo:hi]
    theta = np.exp(np.log(w["theta0"]) + 0.1 * eps)
    x0 = np.tile(w["x0"], (hi - lo, 1))
    # observations: one noisy trajectory at integer times 1..40 (examples/inference.py:58-64)
    ind = np.arange(10, 401, 10)
    Yrng = np.random.default_rng(5)
    k = wl.make_kode(w)
    x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()
    mu_true, _ = k.solve_mv(torch.from_numpy(w["x0"][None]).cuda(), theta=torch.from_numpy(w["theta0"][None]).cuda())
    Y = mu_true[0].cpu().numpy()[ind][:, [0, 3]] + 0.2 * Yrng.standard_normal((40, 2))
    best = 1e30
    for _ in range(a.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ll_local = k.loglik(x0d, thd, Y, ind, [0, 3], 0.2)
        ll = gather_batch(ll_local, a.total)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    if rank == 0:
        print(json.dumps({"config": "C5 fitz loglik sweep", "total": a.total, "n_gpus": world, "ms": best,
                          "solves_per_s": a.total / best * 1e3, "gathered": int(ll.numel()),
                          "argmax_theta_index": int(ll.argmax()), "loglik_checksum": float(ll.sum())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
