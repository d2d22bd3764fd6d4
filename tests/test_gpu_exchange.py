"""GPU tests of the multi-GPU exchange on ONE device: `world` ranks (one api.Context + one Exchange each)
live in this process and reach each other's arenas directly (Exchange.connect_local), so the same
kernels, flags and waits run as between GPUs -- only the stores stay on the device instead of crossing
NVLink.  Checked against the numpy restatement (tests/exchange_model.py) and, for the whole sharded
pipeline, bit for bit against the single-cloud run."""
import numpy as np
import pytest

from exchange_model import exchange_model

pytestmark = pytest.mark.gpu


def _ranks(world, cloud, halo, S_cap=64, node_cap=200000, stride=32):
    import torch
    from polishpathplanning_b200 import api, parallel
    n = cloud.shape[0]
    starts = parallel.index_ranges(n, world)
    ctxs = [api.Context(0) for _ in range(world)]
    cap = int(n * 1.2) + 1024
    exs = [parallel.Exchange(ctxs[r], r, world, starts, cap, S_cap, node_cap, stride) for r in range(world)]
    parallel.Exchange.connect_local(exs)
    dev = torch.device("cuda", 0)
    chunks = [torch.from_numpy(np.ascontiguousarray(cloud[starts[r]:starts[r + 1]])).to(dev) for r in range(world)]
    torch.cuda.synchronize()
    return ctxs, exs, chunks, starts


def _run_exchange(exs, chunks, halo, stride_bytes):
    for p in range(4):                                  # several ranks in one process: phase by phase
        for ex, ch in zip(exs, chunks):
            ex.phase(p, ch.data_ptr(), ch.shape[0], stride_bytes, halo)
    return [ex.finish() for ex in exs]


@pytest.mark.parametrize("world,n,halo", [(2, 50000, 12.0), (3, 60000, 12.0), (5, 40000, 70.0), (8, 30011, 5.0)])
def test_exchange_matches_model(world, n, halo):
    from polishpathplanning_b200 import synth
    cloud = synth.panel(n, seed=31)
    cloud[7, 0] = np.nan
    cloud[n // 2, 2] = np.inf
    cloud[n - 1, 1] = np.nan
    ctxs, exs, chunks, starts = _ranks(world, cloud, halo)
    model, minfo = exchange_model([cloud[starts[r]:starts[r + 1]] for r in range(world)], starts, halo)
    for step in range(2):                               # a second step re-uses flags, tickets and scratch
        infos = _run_exchange(exs, chunks, halo, 32)
        for r in range(world):
            xyz, w = model[r]
            assert infos[r]["n_local"] == xyz.shape[0]
            assert infos[r]["n_owned"] == minfo["n_owned"][r]
            assert np.array_equal(infos[r]["cuts"], minfo["cuts"])
            assert np.array_equal(infos[r]["x_range"], minfo["x_range"])
            slab = ctxs[r].download(exs[r].slab_ptr, (xyz.shape[0], 4), np.float32)
            assert np.array_equal(slab[:, 3].view(np.int32), w)
            assert np.array_equal(slab[:, :3].view(np.uint32), xyz.view(np.uint32))
            rm = ctxs[r].download(exs[r].row_map_ptr, (xyz.shape[0],), np.int32)
            assert np.array_equal(rm, np.where(w >= 0, w, -1))
    assert sum(i["n_owned"] for i in infos) == n
    for ex in exs:
        ex.check()
        ex.close()
    for c in ctxs:
        c.close()


def test_exchange_receive_overflow_is_reported():
    import torch
    from polishpathplanning_b200 import api, parallel, synth
    cloud = synth.panel(20000, seed=2)
    world = 2
    starts = parallel.index_ranges(20000, world)
    ctxs = [api.Context(0) for _ in range(world)]
    exs = [parallel.Exchange(ctxs[r], r, world, starts, 6000, 8, 1000) for r in range(world)]   # slabs hold ~10k + halo
    parallel.Exchange.connect_local(exs)
    chunks = [torch.from_numpy(np.ascontiguousarray(cloud[starts[r]:starts[r + 1]])).to("cuda:0") for r in range(world)]
    torch.cuda.synchronize()
    for p in range(4):
        for ex, ch in zip(exs, chunks):
            ex.phase(p, ch.data_ptr(), ch.shape[0], 32, 12.0)
    for ex in exs:
        with pytest.raises(api.PPPError) as e:
            ex.finish()
        assert e.value.status == -4
    for ex in exs:
        ex.close()
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("world,stride,combined", [(3, 32, False), (4, 16, False), (3, 32, True), (2, 16, True)])
def test_sharded_pipeline_equals_single_cloud(world, stride, combined):
    """Exchange -> per-slab kNN + normals (records delivered to their home rank) -> contours of the owned
    planes (delivered to rank 0's regions) == the same calls on the whole cloud, bit for bit.  combined: the
    one-synchronisation ppp_exch_finish_attach (slab size read from device memory by the ingest kernels) instead of
    ppp_exch_finish + ppp_exch_attach."""
    import torch
    from polishpathplanning_b200 import api, parallel, synth
    n, k, halo, S = 120000, 16, 12.0, 24
    cloud = synth.panel(n, seed=23)
    cloud[1234, 1] = np.nan
    ctxs, exs, chunks, starts = _ranks(world, cloud, halo, S_cap=S, node_cap=60000, stride=stride)
    ref_ctx = api.Context(0)
    full = None
    for step in range(2):
        if step == 1:
            # a DIFFERENT cloud through the same exchange objects: nothing of the previous step (flags, counts,
            # receive buffers, home normals, contour regions) may leak into this one
            cloud = synth.panel(n, seed=29)
            cloud[77, 2] = np.inf
            for r in range(world):
                chunks[r].copy_(torch.from_numpy(np.ascontiguousarray(cloud[starts[r]:starts[r + 1]])))
            torch.cuda.synchronize()
        planes = synth.even_planes(cloud, S)
        if full is not None:
            full.close()
        full = api.Cloud(ref_ctx, cloud)
        ref_n, ref_i = full.normals_knn(k, stride_floats=stride // 4, return_idx=True)
        ro, ry, rx, rz = full.slice_contours(planes, "B")
        if combined:
            for p in range(4):                          # several ranks in one process: phase by phase
                for ex, ch in zip(exs, chunks):
                    ex.phase(p, ch.data_ptr(), ch.shape[0], 32, halo)
            both = [ex.finish_attach(to_rank0=True) for ex in exs]
            infos = [b[0] for b in both]
        else:
            infos = _run_exchange(exs, chunks, halo, 32)
        clouds, pos, idx = [], [], []
        for r in range(world):
            c = both[r][1] if combined else exs[r].attach(to_rank0=True)
            clouds.append(c)
            idx.append(torch.empty((infos[r]["n_local"], k), dtype=torch.int32, device="cuda:0"))
            c.dev_normals_knn(k, exs[r].home_normals_ptr, stride, idx_ptr=idx[r].data_ptr())
            exs[r].results_signal(parallel.Exchange.NORMALS)
            pos.append(parallel.owned_planes(planes, infos[r]["cuts"], r))
            res = c.dev_slice_contours(planes[pos[r]], "B")
            assert res["y"] == exs[0].nodes_region(r)["y"]          # delivered in place, not to the cloud's own buffers
            exs[r].results_signal(parallel.Exchange.CONTOURS)
        for r in range(world):
            exs[r].results_wait(parallel.Exchange.NORMALS)
            exs[r].results_wait(parallel.Exchange.CONTOURS)
            ctxs[r].sync()
        got_n = np.concatenate([exs[r].read_home_normals() for r in range(world)], axis=0)
        assert np.array_equal(got_n.view(np.uint32), ref_n.view(np.uint32))
        per_rank = exs[0].read_nodes([len(p) for p in pos])
        goff, gy, gx, gz = parallel.assemble_contours(S, [(pos[r],) + per_rank[r] for r in range(world)])
        assert np.array_equal(goff, ro) and np.array_equal(gy, ry) and np.array_equal(gx, rx) and np.array_equal(gz, rz)
        # neighbour ids of the owned rows, back in global numbering
        for r in range(world):
            rm = ctxs[r].download(exs[r].row_map_ptr, (infos[r]["n_local"],), np.int32)
            slab_w = ctxs[r].download(exs[r].slab_ptr, (infos[r]["n_local"], 4), np.float32)[:, 3].view(np.int32)
            g_of_row = np.where(slab_w >= 0, slab_w, ~slab_w)
            li = idx[r].cpu().numpy()
            own = rm >= 0
            gi = np.where(li[own] >= 0, g_of_row[np.clip(li[own], 0, None)], -1)
            assert np.array_equal(gi, ref_i[rm[own]])
        for c in clouds:
            c.close()
    for ex in exs:
        ex.check()
        ex.close()
    for c in ctxs:
        c.close()
    full.close()
    ref_ctx.close()
