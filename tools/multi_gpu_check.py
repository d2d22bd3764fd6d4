#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL):
every rank starts with 1/G of the cloud BY INDEX, the records are redistributed into x-slabs +
halo with an NCCL all-to-all-v (parallel.redistribute), each rank runs kNN+normals and the
contours of its planes on its slab, the results are gathered to rank 0 and compared bit for bit
with the single-GPU run of the whole cloud."""
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
    n, k, halo, S = int(os.environ.get("PPP_CHECK_N", "400000")), 16, 12.0, 40
    cloud = synth.panel(n, seed=23)
    planes = synth.even_planes(cloud, S)
    a, b = (n * rank) // world, (n * (rank + 1)) // world
    chunk = torch.from_numpy(cloud[a:b].copy()).to(dev)
    local, g, owned, cuts = parallel.redistribute(dist, chunk, a, rank, world, halo)
    local = local.contiguous()
    ctx = api.Context(lr)
    torch.cuda.synchronize()
    c = api.Cloud(ctx, device_ptr=local.data_ptr(), n=local.shape[0], stride_bytes=32)
    nl = local.shape[0]
    nrm = torch.empty((nl, 4), dtype=torch.float32, device=dev)
    idx = torch.empty((nl, k), dtype=torch.int32, device=dev)
    d2 = torch.empty((nl, k), dtype=torch.float32, device=dev)
    c.dev_normals_knn(k, nrm.data_ptr(), 16, idx_ptr=idx.data_ptr(), d2_ptr=d2.data_ptr())
    ctx.sync()
    bad = parallel.halo_violations(local.cpu().numpy(), owned.cpu().numpy(), d2[:, -1].cpu().numpy(), cuts, rank, halo)
    assert len(bad) == 0, "halo too narrow for %d points" % len(bad)
    pos = parallel.owned_planes(planes, cuts, rank)
    off, y, x, z = c.slice_contours(planes[pos], "B")
    # neighbour ids back to global numbering
    gidx = torch.where(idx >= 0, g[idx.clamp(min=0).to(torch.int64)], torch.full_like(idx, -1, dtype=torch.int64))
    res = parallel.gather_to_rank0(dist, [g[owned], nrm[owned], gidx[owned]], rank, world, device=dev)
    counts = np.diff(off)
    node_plane = np.repeat(pos, counts).astype(np.int64)
    res2 = parallel.gather_to_rank0(dist, [node_plane, y, x, z], rank, world, device=dev)
    # second delivery path: no collective; every rank's kernels store into rank 0's global arrays over
    # NVLink (parallel.PeerSink), two steps to exercise the per-step flags
    smax = torch.tensor([len(pos)], dtype=torch.int64, device=dev)
    dist.all_reduce(smax, op=dist.ReduceOp.MAX)
    sink = parallel.PeerSink(ctx, dist, dev, rank, world, n, node_cap=max(4096, nl // 2), S_cap=int(smax.item()))
    row_map = torch.where(owned, g, torch.full_like(g, -1)).to(torch.int32).contiguous()
    torch.cuda.synchronize()
    c.dev_set_normal_row_map(row_map.data_ptr())
    for step in (1, 2):
        c.dev_normals_knn(k, sink.normals_ptr, 16, idx_ptr=idx.data_ptr())
        sink.attach(c)
        r_dev = c.dev_slice_contours(planes[pos], "B")
        assert r_dev["total_nodes"] <= sink.node_cap
        sink.delivered(step)
    c.dev_set_normal_row_map(None)
    c.dev_set_contour_buffers(None, None, None, 0)
    c.dev_set_contour_offsets_buffer(None, 0)
    ctx.sync()
    all_pos = [None] * world
    dist.all_gather_object(all_pos, pos.tolist())
    ok = True
    if rank == 0:
        peer_n, peer_nodes = sink.read([len(p_) for p_ in all_pos])
        full = api.Cloud(ctx, cloud)
        ref_n, ref_i = full.normals_knn(k, stride_floats=4, return_idx=True)
        got_n = parallel.assemble_normals(n, 4, [(r[0], r[1]) for r in res])
        got_i = np.full((n, k), -2, np.int64)
        for r in res:
            got_i[r[0]] = r[2]
        ro, ry, rx, rz = full.slice_contours(planes, "B")
        per_rank = []
        for r in res2:
            pl, yy, xx, zz = r
            ppos = np.unique(pl)
            o = np.concatenate([[0], np.cumsum([(pl == s).sum() for s in ppos])]).astype(np.int64)
            per_rank.append((ppos.astype(np.int64), o, yy, xx, zz))
        goff, gy, gx, gz = parallel.assemble_contours(S, per_rank)
        poff, py, px, pz = parallel.assemble_contours(
            S, [(np.asarray(all_pos[r], np.int64),) + peer_nodes[r] for r in range(world)])
        peer_ok_n = np.array_equal(peer_n.view(np.uint32), ref_n.view(np.uint32))
        peer_ok_c = np.array_equal(poff, ro) and np.array_equal(py, ry) and np.array_equal(px, rx) and np.array_equal(pz, rz)
        print("MULTI_GPU_CHECK peer_normals=%s peer_contours=%s" % (peer_ok_n, peer_ok_c), flush=True)
        ok = (np.array_equal(got_n.view(np.uint32), ref_n.view(np.uint32)) and np.array_equal(got_i, ref_i.astype(np.int64))
              and np.array_equal(goff, ro) and np.array_equal(gy, ry) and np.array_equal(gz, rz) and peer_ok_n and peer_ok_c)
        print("MULTI_GPU_CHECK world=%d n=%d normals_bitexact=%s knn_ids=%s contours=%s" % (
            world, n, np.array_equal(got_n.view(np.uint32), ref_n.view(np.uint32)), np.array_equal(got_i, ref_i.astype(np.int64)),
            np.array_equal(goff, ro) and np.array_equal(gy, ry) and np.array_equal(gz, rz)), flush=True)
        full.close()
    c.close()
    sink.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
