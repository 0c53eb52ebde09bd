#!/usr/bin/env python
"""How fast can this GPU WRITE?  The backward kernel's traffic is 82 % writes (8.8 GB of outputs against
1.9 GB of history reads); MEASURED_PEAKS.json's HBM figure is a copy (half reads, half writes).  Times
pure fills, a copy and a 4:1 write:read mix with torch ops (library kernels, measurement only)."""
import json
import torch

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

n = 8_800_000_000 // 8
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
small = torch.empty(n // 4, dtype=torch.float64, device="cuda").normal_()
out = {}
ms = timed(lambda: a.fill_(1.0)); out["fill_GBps"] = n * 8 / ms / 1e6
ms = timed(lambda: a.zero_()); out["memset_GBps"] = n * 8 / ms / 1e6
ms = timed(lambda: b.copy_(a)); out["copy_GBps_rw"] = 2 * n * 8 / ms / 1e6
# 4 bytes written per byte read: broadcast-expand a quarter-size source
ms = timed(lambda: torch.add(small.view(1, -1).expand(4, -1), 1.0, out=a.view(4, -1))); out["write4_read1_GBps"] = (n + n // 4) * 8 / ms / 1e6
ms = timed(lambda: torch.sum(a)); out["read_GBps"] = n * 8 / ms / 1e6
print(json.dumps(out))
