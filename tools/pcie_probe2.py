#!/usr/bin/env python
"""Wall time of (1 MB D2H copy + stream sync) on different streams / pinned allocations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from polishpathplanning_b200 import api
dev = torch.device("cuda", 0)
ctx = api.Context(0)
size = 921304
d = torch.empty(size, dtype=torch.uint8, device=dev)
h_torch = torch.empty(size, dtype=torch.uint8).pin_memory()
h_lib = torch.from_numpy(ctx.pinned_empty((size,), np.uint8))
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
streams = {"torch default": torch.cuda.current_stream(), "torch new": torch.cuda.Stream(),
           "torch high prio": torch.cuda.Stream(priority=-1), "lib main (low prio)": torch.cuda.ExternalStream(ctx.stream, device=dev)}
for hname, h in (("torch pinned", h_torch), ("lib pinned", h_lib)):
    for sname, s in streams.items():
        with torch.cuda.stream(s):
            for _ in range(3):
                h.copy_(d, non_blocking=True)
            s.synchronize()
            ts = []
            for _ in range(20):
                t0 = time.perf_counter()
                h.copy_(d, non_blocking=True)
                s.synchronize()
                ts.append(1e6 * (time.perf_counter() - t0))
        ts.sort()
        print("%-14s %-22s median %.1f us  min %.1f" % (hname, sname, ts[10], ts[0]))
