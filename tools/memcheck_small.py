#!/usr/bin/env python
"""A small pass over the hot path for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/memcheck_small.py
k = 8 / 16 / 32 / 64 searches + normals, slicing with one CTA and with clusters of 2 / 4 / 8 CTAs per slice, and the
2-rank exchange on one GPU (what __graft_entry__.smoke() runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from polishpathplanning_b200 import api, synth
ctx = api.Context(0)
cloud = synth.panel(6000, 3)
gc = api.Cloud(ctx, cloud)
for k in (8, 16, 32, 64):
    n = gc.normals_knn(k)
    i, d = gc.knn(k)
    print("k=%d ok, finite normals %.3f" % (k, float(np.isfinite(n[:, 0]).mean())))
planes = synth.even_planes(cloud, 6)
for cl in ("1", "2", "4", "8"):
    os.environ["PPP_SLICE_CLUSTER"] = cl
    off, y, x, z = gc.slice_contours(planes, "B", half_width=6.0)
    print("cluster %s: %d nodes" % (cl, len(y)))
os.environ.pop("PPP_SLICE_CLUSTER")
gc.close()
ctx.close()
import __graft_entry__ as g
g.smoke()
print("done")
