#!/usr/bin/env python
"""Context for the depth kernel's roofline: what a pure-write kernel reaches on this GPU (torch fill_ and a
copy) at the bench's output size (232 MB) and at 4 GB.  MEASURED_PEAKS.json's hbm_gbs is a COPY (read+write)."""
import json

import torch


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
out = {}
for name, nbytes in (("232MB", 232_000_000), ("4GB", 4 << 30)):
    x = torch.empty(nbytes // 4, dtype=torch.int32, device="cuda")
    y = torch.empty_like(x)
    ms = timed(lambda: x.fill_(7))
    out[f"fill_{name}_GBps"] = nbytes / ms / 1e6
    ms = timed(lambda: y.copy_(x))
    out[f"copy_{name}_GBps_read_plus_write"] = 2 * nbytes / ms / 1e6
    del x, y
print(json.dumps(out))
