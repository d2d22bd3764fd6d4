#!/usr/bin/env python
"""How much does the k=16 search slow down when a launch's queries are a 1/C subset (by original
index) of the cloud, cell-sorted within the subset, instead of a contiguous run of the fully sorted
order?  (Feasibility of chunking the search by original row range so that the normals' D2H copy of
one chunk overlaps the search of the next.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from polishpathplanning_b200 import api, synth

n, k = 1_000_000, 16
ctx = api.Context(0)
cloud = synth.panel(n, 0)
c = api.Cloud(ctx, cloud)
c.knn(k)                                   # builds the index
h = 1.35 * np.sqrt(k / np.pi) / 2
cu = np.floor(cloud[:, 0] / h).astype(np.int64)
cv = np.floor(cloud[:, 1] / h).astype(np.int64)
full_order = np.lexsort((cloud[:, 0], cu, cv))          # row-major cell order, like the grid


def kernel_ms(q):
    q = np.ascontiguousarray(q[:, :3])
    for _ in range(2):
        c.knn(k, queries=q)
    ctx.kernel_profile(True); ctx.kernel_profile_read(True)
    for _ in range(5):
        c.knn(k, queries=q)
    prof = ctx.kernel_profile_read(True); ctx.kernel_profile(False)
    return {kk: round(v[0] / 5, 4) for kk, v in prof.items() if v[0] / 5 > 0.002}


for C in (1, 2, 4, 8, 16):
    m = n // C
    coherent = cloud[full_order[:m]]
    rows = np.arange(m)                                    # chunk 0 by original index
    sub = rows[np.lexsort((cloud[rows, 0], cu[rows], cv[rows]))]
    print("C=%2d  %7d queries | contiguous run of sorted order: %s | 1/C subset, cell-sorted: %s"
          % (C, m, kernel_ms(coherent), kernel_ms(cloud[sub])), flush=True)
