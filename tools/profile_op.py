#!/usr/bin/env python
"""Per-kernel CUDA-event times of one operation (kernels serialised).  python tools/profile_op.py radius|knn16|knn32|contoursA|contoursB [points] [slices]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from polishpathplanning_b200 import api, synth
op = sys.argv[1] if len(sys.argv) > 1 else "radius"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ctx = api.Context(0); dev = torch.device("cuda", 0)
cloud = synth.panel(n, 0); raw = torch.from_numpy(cloud).to(dev)
nrm = torch.empty((n, 8), dtype=torch.float32, device=dev); idx = torch.empty((n, 64), dtype=torch.int32, device=dev)
S = int(sys.argv[3]) if len(sys.argv) > 3 else 200
planes = synth.even_planes(cloud, S)
def step():
    c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
    if op == "radius": c.dev_normals_radius(2.5, nrm.data_ptr(), 32)
    elif op.startswith("knn"): c.dev_normals_knn(int(op[3:]), nrm.data_ptr(), 32, idx_ptr=idx.data_ptr())
    elif op == "contoursA": c.dev_index(16, 0); c.dev_slice_contours(planes, "A")
    elif op == "contoursB": c.dev_index(16, 0); c.dev_slice_contours(planes, "B")
    c.close()
for _ in range(3): step()
ctx.sync(); ctx.kernel_profile(True); ctx.kernel_profile_read(True)
for _ in range(5): step()
prof = ctx.kernel_profile_read(True)
tot = 0
for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("%-24s %8.4f ms/step  (%d launches/step)" % (k, ms / 5, cnt // 5)); tot += ms / 5
print("sum %.4f ms" % tot)
