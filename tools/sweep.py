#!/usr/bin/env python
"""Size / k / radius / slice-count sweep of the device-resident path on one GPU (BASELINE.json
configs 3-5 shapes on a single B200).  Prints a table; every row is also checked for basic sanity.
    python tools/sweep.py [--quick]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from polishpathplanning_b200 import api, synth  # noqa: E402


def timed(ctx, fn, reps=3):
    fn()
    ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.timer_read(1, reset=True)
        ctx.timer_begin(1)
        fn()
        ctx.timer_end(1)
        ms, _ = ctx.timer_read(1, reset=True)
        best = min(best, ms)
    return best


def main():
    quick = "--quick" in sys.argv
    ctx = api.Context(0)
    dev = torch.device("cuda", 0)
    rows = []
    sizes = [1_000_000, 5_000_000] if quick else [1_000_000, 5_000_000, 10_000_000, 20_000_000]
    for n in sizes:
        cloud = synth.panel(n, seed=0)
        raw = torch.from_numpy(cloud).to(dev)
        nrm = torch.empty((n, 8), dtype=torch.float32, device=dev)
        for k in (8, 16, 32, 64):
            idx = torch.empty((n, k), dtype=torch.int32, device=dev)

            def step():
                c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
                c.dev_normals_knn(k, nrm.data_ptr(), 32, idx_ptr=idx.data_ptr())
                c.close()
            ms = timed(ctx, step)
            ok = bool((idx[:, 0] == torch.arange(n, device=dev, dtype=torch.int32)).all().item())
            rows.append(("knn+normals", n, "k=%d" % k, ms, n / ms * 1e-6, n * (32 + 4 * k) / ms * 1e-6, ok))
            del idx

        def stepr():
            c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
            c.dev_normals_radius(2.5, nrm.data_ptr(), 32)
            c.close()
        ms = timed(ctx, stepr)
        ok = bool(torch.isfinite(nrm[:, 0]).float().mean().item() > 0.999)
        rows.append(("radius normals", n, "r=2.5", ms, n / ms * 1e-6, n * (32 + 4 * 19.6) / ms * 1e-6, ok))
        c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
        c.dev_index(16, 0.0)
        for S in ((50, 200, 1000) if quick else (50, 200, 1000, 5000)):
            planes = synth.even_planes(cloud, S)
            res = {}

            def steps():
                res.update(c.dev_slice_contours(planes, "B"))
            ms = timed(ctx, steps)
            rows.append(("bands+contours B", n, "S=%d" % S, ms, n / ms * 1e-6,
                         (16 * n + 20 * res["total_members"] + 24 * res["total_nodes"]) / ms * 1e-6, res["total_nodes"] > 0))
        c.close()
        del raw, nrm
        torch.cuda.empty_cache()
    print("%-18s %10s %8s %10s %12s %10s %s" % ("stage", "points", "param", "ms", "Gpts/s", "alg GB/s", "ok"))
    for r in rows:
        print("%-18s %10d %8s %10.3f %12.1f %10.1f %s" % r)
    ctx.close()


if __name__ == "__main__":
    main()
