#!/usr/bin/env python
"""Context for the generate kernel's roofline fraction: what a WRITE-ONLY stream reaches on this GPU next to the copy figure of
MEASURED_PEAKS.json (read + write bytes per second).  The generate kernel writes 104 B per path and reads nothing.  Library kernels
(torch fill / copy), used as a yardstick only."""
import json

import torch

n = 4 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


t_fill = timed(lambda: a.zero_())
t_copy = timed(lambda: b.copy_(a))
print(json.dumps({"fill_GBps": n / t_fill / 1e9, "copy_GBps_read_plus_write": 2 * n / t_copy / 1e9, "bytes": n}))
