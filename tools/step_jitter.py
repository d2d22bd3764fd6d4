#!/usr/bin/env python
"""Per-step device time of the cfg2 step, one CUDA-event region per step: distribution and outliers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from polishpathplanning_b200 import api, synth
n = 1_000_000
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = api.Context(0); dev = torch.device("cuda", 0)
cloud = synth.panel(n, 0); raw = torch.from_numpy(cloud).to(dev)
nrm = torch.empty((n, 8), dtype=torch.float32, device=dev); idx = torch.empty((n, 16), dtype=torch.int32, device=dev)
planes = synth.even_planes(cloud, 200)
def step():
    c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
    c.dev_normals_knn(16, nrm.data_ptr(), 32, idx_ptr=idx.data_ptr())
    c.dev_slice_contours(planes, "B")
    c.close()
for _ in range(5): step()
ctx.sync()
ms = []
for i in range(steps):
    ctx.timer_begin(3); step(); ctx.timer_end(3)
    t, k = ctx.timer_read(3, True)
    ms.append(t)
ms = np.array(ms)
print("steps %d: median %.4f  mean %.4f  p90 %.4f  max %.4f ms" % (steps, np.median(ms), ms.mean(), np.percentile(ms, 90), ms.max()))
out = np.flatnonzero(ms > 3 * np.median(ms))
print("outliers (> 3x median):", [(int(i), round(float(ms[i]), 3)) for i in out[:20]])
