#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
every rank starts with 1/G of the cloud BY INDEX on its GPU, the records are redistributed into x-slabs +
halo by the exchange kernels over NVLink (parallel.Exchange over CUDA IPC), each rank runs kNN + normals
and the contours of its planes on its slab, normal records travel to their home rank and contour nodes to
rank 0, and the assembled result is compared bit for bit with the single-GPU run of the whole cloud.
Two steps, so that the per-step flags and the re-use of every buffer are exercised; the cloud changes
between the steps so that stale data cannot pass."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from polishpathplanning_b200 import api, parallel, synth  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    n, k, halo, S = int(os.environ.get("PPP_CHECK_N", "400000")), int(os.environ.get("PPP_CHECK_K", "16")), 12.0, 40
    ctx = api.Context(lr)
    starts = parallel.index_ranges(n, world)
    a, b = int(starts[rank]), int(starts[rank + 1])
    ex = parallel.Exchange.over_dist(ctx, dist, dev, rank, world, n, int((b - a) * 1.4) + 4096, S, max(4096, n // 2), 16)
    ok = True
    for step, seed in enumerate((23, 24)):
        cloud = synth.panel(n, seed=seed)
        cloud[17 + step, 1] = np.nan
        planes = synth.even_planes(cloud, S)
        chunk = torch.from_numpy(cloud[a:b].copy()).to(dev)
        torch.cuda.synchronize()
        ex.exchange(chunk.data_ptr(), b - a, 32, halo)
        info, c = ex.finish_attach(to_rank0=True)
        nl = info["n_local"]
        idx = torch.empty((nl, k), dtype=torch.int32, device=dev)
        d2 = torch.empty((nl, k), dtype=torch.float32, device=dev)
        c.dev_normals_knn(k, ex.home_normals_ptr, 16, idx_ptr=idx.data_ptr(), d2_ptr=d2.data_ptr())
        ex.results_signal(ex.NORMALS)
        pos = parallel.owned_planes(planes, info["cuts"], rank)
        res = c.dev_slice_contours(planes[pos], "B")
        assert res["total_nodes"] <= ex.node_cap
        ex.results_signal(ex.CONTOURS)
        ex.results_wait(ex.NORMALS)
        ex.results_wait(ex.CONTOURS)
        ctx.sync()
        ex.check()
        # halo wide enough for every owned point?
        slab = ctx.download(ex.slab_ptr, (nl, 4), np.float32)
        owned = slab[:, 3].view(np.int32) >= 0
        bad = parallel.halo_violations(slab, owned, d2[:, -1].cpu().numpy(), info["cuts"], rank, halo)
        assert len(bad) == 0, "halo too narrow for %d points" % len(bad)
        home = ex.read_home_normals()
        # neighbour ids of the owned rows in global numbering
        w = slab[:, 3].view(np.int32)
        g_of_row = np.where(w >= 0, w, ~w).astype(np.int64)
        li = idx.cpu().numpy()
        gi = np.where(li >= 0, g_of_row[np.clip(li, 0, None)], -1)
        res1 = parallel.gather_to_rank0(dist, [g_of_row[owned], gi[owned]], rank, world, device=dev)
        res2 = parallel.gather_to_rank0(dist, [home], rank, world, device=dev)
        all_pos = [None] * world
        dist.all_gather_object(all_pos, pos.tolist())
        if rank == 0:
            full = api.Cloud(ctx, cloud)
            ref_n, ref_i = full.normals_knn(k, stride_floats=4, return_idx=True)
            ro, ry, rx, rz = full.slice_contours(planes, "B")
            full.close()
            got_n = np.concatenate([r[0] for r in res2], axis=0)
            got_i = np.full((n, k), -2, np.int64)
            for r in res1:
                got_i[r[0]] = r[1]
            per_rank = ex.read_nodes([len(p_) for p_ in all_pos])
            goff, gy, gx, gz = parallel.assemble_contours(S, [(np.asarray(all_pos[r], np.int64),) + per_rank[r] for r in range(world)])
            ok_n = np.array_equal(got_n.view(np.uint32), ref_n.view(np.uint32))
            ok_i = np.array_equal(got_i, ref_i.astype(np.int64))
            ok_c = np.array_equal(goff, ro) and np.array_equal(gy, ry) and np.array_equal(gx, rx) and np.array_equal(gz, rz)
            print("MULTI_GPU_CHECK world=%d n=%d k=%d step=%d normals_bitexact=%s knn_ids=%s contours=%s" % (
                world, n, k, step, ok_n, ok_i, ok_c), flush=True)
            ok = ok and ok_n and ok_i and ok_c
        c.close()
        del chunk, idx, d2
    ex.close(dist)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
