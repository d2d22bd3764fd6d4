#!/usr/bin/env python
"""BASELINE.json configs[0]: ~100k-point workpiece through the ./main flow (estimate_normal r=2.5 +
Contact_Path_Generation sweep R=15, gen-2 pairing).  Times the C++ adapter binary on the GPU and the
CPU oracle (1 thread = what the reference does, and all threads) on the same PCD."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import ppp_oracle as po
from polishpathplanning_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d = tempfile.mkdtemp()
pcd = os.path.join(d, "workpiece.pcd")
synth.write_pcd(pcd, synth.to_pointxyzrgb(synth.panel_metres(n, 0)))
exe = os.path.join(ROOT, "polishpathplanning_b200", "host", "ppp_main")
for rep in range(2):
    t0 = time.perf_counter()
    r = subprocess.run([exe, pcd], capture_output=True, text=True, cwd=d)
    wall = time.perf_counter() - t0
print("ppp_main (2nd run): wall %.3f s incl. process start, CUDA init, PCD read; sweep line: %s" % (
    wall, [l for l in r.stdout.splitlines() if "Using Time" in l]))
cloud = synth.panel(n, 0)
for threads in (1, 0):
    t0 = time.perf_counter()
    oc = po.OracleCloud(cloud)
    t1 = time.perf_counter()
    oc.normals(radius=2.5, threads=threads)
    t2 = time.perf_counter()
    mn, mx = oc.minmax()
    planes = po.planes("gen2_contact", mn[0], mx[0], 15.0)
    oc.slice_contours(planes, "A", threads=threads)
    t3 = time.perf_counter()
    print("oracle threads=%s: kd-tree %.3f s, normals %.3f s, sweep(%d planes, variant A) %.3f s, total %.3f s" % (
        threads or po.num_threads(), t1 - t0, t2 - t1, len(planes), t3 - t2, t3 - t0))
# library-only timing of the same flow (no process start / file I/O)
from polishpathplanning_b200 import api, reference_api as ra
ctx = api.Context(0)
for rep in range(3):
    t0 = time.perf_counter()
    pg = ra.path_generater(cloud.copy(), 15, ctx=ctx)
    pg.cloud[:, :3] = cloud[:, :3]
    t1 = time.perf_counter()
    pg.estimate_normal()
    t2 = time.perf_counter()
    pg.Contact_Path_Generation()
    t3 = time.perf_counter()
    pg._invalidate()
print("GPU via C ABI (host buffers): upload+normals %.2f ms, sweep %.2f ms" % ((t2 - t1) * 1e3, (t3 - t2) * 1e3))
