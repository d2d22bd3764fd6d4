#!/usr/bin/env python
"""Wall-clock breakdown of the end-to-end (host-pointer) step of bench.py on one GPU:
upload+ingest, bbox, normals-only, contours-only, and the combined call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from polishpathplanning_b200 import api, synth

n = 1_000_000
ctx = api.Context(0)
cloud = synth.panel(n, 0)
planes = synth.even_planes(cloud, 200)
pin_cloud = ctx.pinned_empty(cloud.shape, np.float32); pin_cloud[...] = cloud
pin_normals = ctx.pinned_empty((n, 8), np.float32)
pin_nodes = tuple(ctx.pinned_empty((200000,), np.float64) for _ in range(3))


def timed(name, fn, reps=20):
    for _ in range(3):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    print("%-34s %.3f ms" % (name, 1e3 * (time.perf_counter() - t0) / reps))


def upload():
    c = api.Cloud(ctx, pin_cloud); c.close()


def upload_bbox_index():
    c = api.Cloud(ctx, pin_cloud); c.bbox(); c.dev_index(16, 0); ctx.sync(); c.close()


keep = api.Cloud(ctx, pin_cloud)
timed("upload + ingest (Cloud ctor)", upload)
timed("upload + bbox + index", upload_bbox_index)
timed("normals_knn only (resident cloud)", lambda: keep.normals_knn(16, out=pin_normals) if hasattr(keep, "normals_knn") else None)
timed("slice_contours only (resident)", lambda: keep.slice_contours(planes, "B", out=pin_nodes))
timed("normals_and_contours (resident)", lambda: keep.normals_and_contours(planes, "B", k=16, normals_out=pin_normals, nodes_out=pin_nodes))


def full():
    c = api.Cloud(ctx, pin_cloud)
    c.bbox()
    c.normals_and_contours(planes, "B", k=16, normals_out=pin_normals, nodes_out=pin_nodes)
    c.close()


timed("full e2e step", full)
keep.close()
