#!/usr/bin/env python
"""BASELINE.json configs[3] and [4] on N GPUs (run under torchrun for N > 1, plain python for N = 1):
  cfg4  5M-point panel, k = 16 normals, slice-count sweep 50 .. 5000
  cfg5  20M-point panel, kNN k sweep 8 .. 64 and the reference's radius-2.5 mode, 200 slices
Device-resident step (exchange included for N > 1), CUDA events, max over ranks; one line per point.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_multi.py [cfg4|cfg5|all]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class Args:
    gpus = int(os.environ.get("WORLD_SIZE", "1"))


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    env = bench.Env(Args)
    measure = bench.measure_single if env.world == 1 else bench.measure_multi
    points = []
    if which in ("cfg4", "all"):
        for S in (50, 100, 200, 500, 1000, 2000, 5000):
            points.append(("cfg4", {"name": "cfg4_S%d" % S, "n_total": 5_000_000, "k": 16, "S": S, "gpus": env.world}))
    if which in ("cfg5", "all"):
        for k in (8, 16, 32, 64):
            points.append(("cfg5", {"name": "cfg5_k%d" % k, "n_total": 20_000_000, "k": k, "S": 200, "gpus": env.world}))
        points.append(("cfg5", {"name": "cfg5_r2p5", "n_total": 20_000_000, "k": 16, "radius": 2.5, "S": 200, "gpus": env.world}))
    for fam, cfg in points:
        kw = dict(e2e=False) if env.world == 1 else dict(verify=False, e2e=False)
        m = measure(env, cfg, 3, 3, **kw)
        if env.rank == 0:
            top = sorted(m["prof"].items(), key=lambda kv: -kv[1][0])[:3]
            print(json.dumps({"config": fam, "name": cfg["name"], "n_gpus": env.world, "points": cfg["n_total"], "k": cfg["k"],
                              "radius": cfg.get("radius"), "slices": cfg["S"], "ms_per_step": m["total_ms"] / 3,
                              "points_per_s": cfg["n_total"] * 3 / (m["total_ms"] * 1e-3), "exchange_ms": m["exchange_ms"],
                              "top_kernels_ms": {kk: round(v[0] / 3, 4) for kk, v in top}}), flush=True)
    env.close()


if __name__ == "__main__":
    main()
