#!/usr/bin/env python
"""Probe: does running the headline batch as TWO half batches on two streams (forward of one half overlapping the
HBM-bound smoother of the other) beat one launch pair?  Two handles, two torch streams; prints ms per 65,536 systems."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rodeo_legacy_b200 import workloads as wl

B = 65536
w, x0, theta = wl.fitz_batch(B)
x0d, thd = torch.from_numpy(x0).cuda(), torch.from_numpy(theta).cuda()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


k = wl.make_kode(w)
print("one launch pair: %.3f ms" % timed(lambda: k.solve_mv(x0d, theta=thd)))
for parts in (2, 4):
    ks = [wl.make_kode(w) for _ in range(parts)]
    ss = [torch.cuda.Stream() for _ in range(parts)]
    h = B // parts

    def run():
        cur = torch.cuda.current_stream()
        for i in range(parts):
            ss[i].wait_stream(cur)
            with torch.cuda.stream(ss[i]):
                ks[i].solve_mv(x0d[i * h:(i + 1) * h], theta=thd[i * h:(i + 1) * h])
        for i in range(parts):
            cur.wait_stream(ss[i])

    print("%d parts on %d streams: %.3f ms" % (parts, parts, timed(run)))
