#!/usr/bin/env python
"""Timeline of ONE device-resident cfg2 step: start/end of every launch on its own stream (overlap kept), and the
idle gaps of the main stream.  python tools/step_timeline.py [points] [k] [slices]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from polishpathplanning_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 200
ctx = api.Context(0); dev = torch.device("cuda", 0)
cloud = synth.panel(n, 0); raw = torch.from_numpy(cloud).to(dev)
nrm = torch.empty((n, 8), dtype=torch.float32, device=dev); idx = torch.empty((n, k), dtype=torch.int32, device=dev)
planes = synth.even_planes(cloud, S)
def step():
    c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
    c.dev_normals_knn(k, nrm.data_ptr(), 32, idx_ptr=idx.data_ptr())
    c.dev_slice_contours(planes, "B")
    c.close()
for _ in range(5): step()
ctx.sync()
best = None
for rep in range(5):
    ctx.kernel_trace(True); step(); ctx.sync(); tr = ctx.kernel_trace_read(); ctx.kernel_trace(False)
    span = max(t[3] for t in tr) - min(t[2] for t in tr)
    if best is None or span < best[0]: best = (span, tr)
span, tr = best
t00 = min(t[2] for t in tr)
print("step span %.4f ms, %d launches (best of 5, events around every launch add a little)" % (span, len(tr)))
busy = 0.0
for name, aux, t0, t1 in tr:
    print("%-18s %s  %8.4f -> %8.4f  (%.4f)" % (name, "aux " if aux else "main", t0 - t00, t1 - t00, t1 - t0))
iv = sorted((t[2], t[3]) for t in tr)
cov, cur0, cur1 = 0.0, iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > cur1: cov += cur1 - cur0; cur0, cur1 = a, b
    else: cur1 = max(cur1, b)
cov += cur1 - cur0
print("device busy (union of launches) %.4f ms, idle inside the step %.4f ms" % (cov, span - cov))
