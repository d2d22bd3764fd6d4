#!/usr/bin/env python
"""H2D / D2H rate of 32 MB copies: cudaHostAlloc memory vs a registered shared-memory mapping (SharedHost)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from polishpathplanning_b200 import api, parallel
ctx = api.Context(0)
n = 32 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
pin = ctx.pinned_empty((n,), np.uint8); pin[:] = 1
sh = parallel.SharedHost(None, 0, 1, n + 4096, "probe", ctx)
sh.array(np.uint8, (n,), 4096)[:] = 2
def rate(fn, reps=20):
    for _ in range(3): fn()
    ctx.sync(); t = time.perf_counter()
    for _ in range(reps): fn()
    ctx.sync(); return n * reps / (time.perf_counter() - t) / 1e9
print("h2d hostalloc  %.1f GB/s" % rate(lambda: ctx.upload(d.data_ptr(), pin.ctypes.data, n)))
print("d2h hostalloc  %.1f GB/s" % rate(lambda: ctx.download_async(pin.ctypes.data, d.data_ptr(), n)))
print("h2d registered %.1f GB/s" % rate(lambda: ctx.upload(d.data_ptr(), sh.host_base + 4096, n)))
print("d2h registered %.1f GB/s" % rate(lambda: ctx.download_async(sh.host_base + 4096, d.data_ptr(), n)))
print("d2h registered via device alias %.1f GB/s" % rate(lambda: ctx.download_async(sh.dev_base + 4096, d.data_ptr(), n)))
