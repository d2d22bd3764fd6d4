#!/usr/bin/env python
"""Where does ppp_knn (self queries) differ from the oracle?  python tools/debug_knn.py [k] [n] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import ppp_oracle as po
from polishpathplanning_b200 import api, synth
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = api.Context(0)
cloud = synth.panel(n, seed)
oc = po.OracleCloud(cloud)
gc = api.Cloud(ctx, cloud)
gi, gd = gc.knn(k)
oi, od = oc.knn(k, threads=0)
bad_i = np.flatnonzero((gi != oi).any(axis=1))
bad_d = np.flatnonzero((gd.view(np.uint32) != od.view(np.uint32)).any(axis=1))
print("k=%d n=%d: rows with idx mismatch %d, d2 mismatch %d" % (k, n, len(bad_i), len(bad_d)))
for r in list(bad_d[:4]) + list(bad_i[:4]):
    print("row", r, "point", cloud[r, :3])
    print("  gpu idx", gi[r]); print("  cpu idx", oi[r]); print("  gpu d2 ", gd[r]); print("  cpu d2 ", od[r])
