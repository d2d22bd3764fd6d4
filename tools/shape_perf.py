#!/usr/bin/env python
"""Device-resident step time (ingest + index + kNN(k=16) normals + 200 slices pairing B) on the panel and on the
non-height-field workpieces, 1M points each.   python tools/shape_perf.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from polishpathplanning_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = api.Context(0); dev = torch.device("cuda", 0)
for name, cloud in (("panel", synth.panel(n, 0)), ("cylinder", synth.cylinder(n, 0)), ("box_with_walls", synth.box_with_walls(n, 0))):
    raw = torch.from_numpy(cloud).to(dev)
    nrm = torch.empty((n, 8), dtype=torch.float32, device=dev); idx = torch.empty((n, 16), dtype=torch.int32, device=dev)
    planes = synth.even_planes(cloud, 200)
    def step():
        c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
        c.dev_normals_knn(16, nrm.data_ptr(), 32, idx_ptr=idx.data_ptr())
        r = c.dev_slice_contours(planes, "B")
        c.close()
        return r
    for _ in range(3): r = step()
    ctx.sync(); ctx.timer_read(2, True)
    for _ in range(10):
        ctx.timer_begin(2); step(); ctx.timer_end(2)
    ms, k = ctx.timer_read(2, True)
    ctx.kernel_profile(True); ctx.kernel_profile_read(True)
    for _ in range(3): step()
    prof = ctx.kernel_profile_read(True); ctx.kernel_profile(False)
    top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:4]
    print("%-16s %8.3f ms/step  members %d nodes %d   %s" % (name, ms / k, r["total_members"], r["total_nodes"],
          ", ".join("%s %.3f" % (kk, v[0] / 3) for kk, v in top)), flush=True)
