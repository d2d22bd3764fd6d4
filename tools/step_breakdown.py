#!/usr/bin/env python
"""Device-resident step time with parts of the path disabled (how much the two streams overlap)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from polishpathplanning_b200 import api, synth
n = 1_000_000
ctx = api.Context(0); dev = torch.device("cuda", 0)
cloud = synth.panel(n, 0); raw = torch.from_numpy(cloud).to(dev)
nrm = torch.empty((n, 8), dtype=torch.float32, device=dev); idx = torch.empty((n, 16), dtype=torch.int32, device=dev)
planes = synth.even_planes(cloud, 200)
def step(knn, sl):
    c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
    if knn: c.dev_normals_knn(16, nrm.data_ptr(), 32, idx_ptr=idx.data_ptr())
    else: c.dev_index(16, 0)
    if sl: c.dev_slice_contours(planes, "B")
    c.close()
for name, knn, sl in (("ingest+index only", 0, 0), ("+ kNN/normals", 1, 0), ("+ slicing", 0, 1), ("everything", 1, 1)):
    for _ in range(3): step(knn, sl)
    ctx.sync(); ctx.timer_read(2, True)
    for _ in range(20):
        ctx.timer_begin(2); step(knn, sl); ctx.timer_end(2)
    ms, k = ctx.timer_read(2, True)
    print("%-22s %.4f ms" % (name, ms / k))
