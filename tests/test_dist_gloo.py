"""N>1 host logic on CPU: two gloo ranks shard a batch, each evaluates its slice (with the
CPU oracle standing in for the GPU kernels -- tests only), one all_gather assembles the result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.common import load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from rodeo_legacy_b200.dist import sharded_loglik, shard_range, gather_batch
        g = load_golden("fitz_n400")
        p = oracle.Problem("fitz", g["W"], 0.0, 40.0, 400, g["Q"], g["c"], g["R"])
        rng = np.random.default_rng(0)
        theta = np.exp(np.log([0.2, 0.2, 3.0]) + 0.1 * rng.standard_normal((batch, 3)))
        x0 = np.tile(g["x0"], (batch, 1))
        ind, Y, gam = g["thin_ind"], g["Y"], float(g["gamma"])

        def local(x0s, ths):
            r = oracle.solve_batch(p, x0s.numpy(), ths.numpy(), None, oracle.MODE_MV, nthreads=2)
            return torch.tensor([oracle.loglik(m, ind, [0, 3], Y, gam) for m in r["mu"]],
                                dtype=torch.float64)

        full = sharded_loglik(local, torch.from_numpy(x0), torch.from_numpy(theta))
        lo, hi = shard_range(batch, rank, world)
        # a ragged 2-D gather as well
        g2 = gather_batch(torch.arange(lo, hi, dtype=torch.float64)[:, None].repeat(1, 3), batch)
        if rank == 0:
            q.put((full.numpy(), g2.numpy(), theta, x0))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_loglik_matches_single():
    import oracle
    batch, world = 7, 2          # ragged on purpose: shards of 4 and 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    full, g2, theta, x0 = q.get(timeout=180)
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    g = load_golden("fitz_n400")
    p = oracle.Problem("fitz", g["W"], 0.0, 40.0, 400, g["Q"], g["c"], g["R"])
    r = oracle.solve_batch(p, x0, theta, None, oracle.MODE_MV)
    want = [oracle.loglik(m, g["thin_ind"], [0, 3], g["Y"], float(g["gamma"])) for m in r["mu"]]
    np.testing.assert_allclose(full, want, rtol=1e-13)
    np.testing.assert_array_equal(g2[:, 0], np.arange(batch))
