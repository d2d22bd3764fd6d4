"""CPU tests of the host-side logic: synthetic clouds, PCD I/O, plane sweeps, and the multi-GPU
sharding (x-slabs + halo + gather) driven by the oracle on CPU, incl. a world_size-2 gloo run."""
import os
import sys

import numpy as np
import pytest

from oracle import ppp_oracle as po
from polishpathplanning_b200 import parallel, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_panel_deterministic_and_layout():
    a = synth.panel(5000, 3)
    b = synth.panel(5000, 3)
    assert a.dtype == np.float32 and a.shape == (5000, 8) and np.array_equal(a, b)
    assert not np.array_equal(a, synth.panel(5000, 4))
    assert np.all(a[:, 3] == 1.0)
    assert np.all(a[:, 4].view(np.uint32) & 0xFFFFFF == 0xFFFFFF)     # r = g = b = 255
    m = synth.panel_metres(5000, 3)
    assert np.array_equal(a[:, :3], m * np.float32(1000.0))              # the reference's *= 1000
    L = np.sqrt(5000)
    assert a[:, 0].min() >= 0 and a[:, 0].max() < L * 1.001


@pytest.mark.parametrize("binary", [True, False])
def test_pcd_roundtrip(tmp_path, binary):
    c = synth.to_pointxyzrgb(synth.panel_metres(300, 2), rgb=0x00102030)
    p = str(tmp_path / "w.pcd")
    synth.write_pcd(p, c, binary=binary)
    r = synth.read_pcd(p)
    assert r.shape == c.shape
    assert np.array_equal(r[:, :3], c[:, :3])
    if binary:
        assert np.array_equal(r[:, 4].view(np.uint32), c[:, 4].view(np.uint32))


def _oracle_rank(cloud_g, planes_g, cuts, rank, halo, k, mode):
    local_idx, owned = parallel.slab_select(cloud_g, cuts, rank, halo)
    local = np.ascontiguousarray(cloud_g[local_idx])
    oc = po.OracleCloud(local)
    nrm, _ = oc.normals(k=k)
    idx, d2 = oc.knn(k)
    bad = parallel.halo_violations(local, owned, d2[:, -1], cuts, rank, halo)
    pos = parallel.owned_planes(planes_g, cuts, rank)
    off, y, x, z = oc.slice_contours(planes_g[pos], mode)
    return local_idx[owned], nrm[owned], bad, (pos, off, y, x, z)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_sharding_equals_single(world):
    cloud = synth.panel(30000, 9)
    planes = synth.even_planes(cloud, 12)
    k, halo = 16, 12.0
    oc = po.OracleCloud(cloud)
    ref_n, _ = oc.normals(k=k)
    ref_c = oc.slice_contours(planes, "B")
    cuts = parallel.slab_cuts(cloud[:, 0], world)
    parts = [_oracle_rank(cloud, planes, cuts, r, halo, k, "B") for r in range(world)]
    owned_total = sum(len(p[0]) for p in parts)
    assert owned_total == cloud.shape[0]                               # every point owned exactly once
    assert all(len(p[2]) == 0 for p in parts)                          # halo wide enough
    full = parallel.assemble_normals(cloud.shape[0], 4, [(p[0], p[1]) for p in parts])
    assert np.array_equal(full.view(np.uint32), ref_n.view(np.uint32))
    goff, y, x, z = parallel.assemble_contours(len(planes), [p[3] for p in parts])
    assert np.array_equal(goff, ref_c[0]) and np.array_equal(y, ref_c[1]) and np.array_equal(z, ref_c[3])


def test_halo_violation_detected():
    cloud = synth.panel(20000, 9)
    cuts = parallel.slab_cuts(cloud[:, 0], 2)
    local_idx, owned = parallel.slab_select(cloud, cuts, 0, 0.5)       # far too narrow
    local = np.ascontiguousarray(cloud[local_idx])
    _, d2 = po.OracleCloud(local).knn(16)
    assert len(parallel.halo_violations(local, owned, d2[:, -1], cuts, 0, 0.5)) > 0


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cloud = synth.panel(20000, 5)
    planes = synth.even_planes(cloud, 8)
    cuts = parallel.slab_cuts(cloud[:, 0], world)
    gidx, nrm, bad, (pos, off, y, x, z) = _oracle_rank(cloud, planes, cuts, rank, 12.0, 16, "B")
    assert len(bad) == 0
    g1 = parallel.gather_to_rank0(dist, [gidx, nrm], rank, world)
    counts = np.diff(off)
    node_plane = np.repeat(pos, counts)
    g2 = parallel.gather_to_rank0(dist, [node_plane, y, x, z], rank, world)
    if rank == 0:
        full = parallel.assemble_normals(cloud.shape[0], 4, [(g[0], g[1]) for g in g1])
        per_rank = []
        for g in g2:
            pl, yy, xx, zz = g
            ppos = np.unique(pl)
            o = np.concatenate([[0], np.cumsum([(pl == s).sum() for s in ppos])]).astype(np.int64)
            per_rank.append((ppos.astype(np.int64), o, yy, xx, zz))
        goff, gy, gx, gz = parallel.assemble_contours(len(planes), per_rank)
        np.savez(os.path.join(out_dir, "gathered.npz"), normals=full, off=goff, y=gy, z=gz)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(str(tmp_path / "gathered.npz"))
    cloud = synth.panel(20000, 5)
    planes = synth.even_planes(cloud, 8)
    oc = po.OracleCloud(cloud)
    ref_n, _ = oc.normals(k=16)
    ref_c = oc.slice_contours(planes, "B")
    assert np.array_equal(got["normals"].view(np.uint32), ref_n.view(np.uint32))
    assert np.array_equal(got["off"], ref_c[0]) and np.array_equal(got["y"], ref_c[1]) and np.array_equal(got["z"], ref_c[3])


def test_exchange_model_is_a_slab_partition():
    """The restated exchange (tests/exchange_model.py, what the GPU kernels are checked against): every
    point owned exactly once, slabs in ascending global index, halo copies exactly the points within
    `halo` of the slab, equal-count cuts."""
    from exchange_model import exchange_model
    cloud = synth.panel(40000, 4)
    cloud[11, 0] = np.nan
    for world, halo in ((2, 12.0), (5, 12.0), (4, 90.0)):
        starts = parallel.index_ranges(cloud.shape[0], world)
        slabs, info = exchange_model([cloud[starts[r]:starts[r + 1]] for r in range(world)], starts, halo)
        cuts = info["cuts"]
        owned_all = np.concatenate([w[w >= 0] for _, w in slabs])
        assert np.array_equal(np.sort(owned_all), np.arange(cloud.shape[0]))
        fin = np.isfinite(cloud[:, :3]).all(axis=1)
        counts = [int(((w >= 0)).sum()) for _, w in slabs]
        assert max(counts) - min(counts) <= cloud.shape[0] // 100 + 2          # 1024 bins: within 1 %
        for r, (xyz, w) in enumerate(slabs):
            g = np.where(w >= 0, w, ~w)
            assert np.all(np.diff(g) > 0)
            assert np.array_equal(xyz.view(np.uint32), cloud[g, :3].view(np.uint32))
            x = cloud[:, 0].astype(np.float64)
            with np.errstate(invalid="ignore"):
                want = fin & (x >= cuts[r] - halo) & (x < cuts[r + 1] + halo)
            want[~fin] = r == 0
            # membership may differ from the plain x-interval only by bin rounding at the two owner limits
            diff = np.setxor1d(np.nonzero(want)[0], g)
            assert len(diff) <= 2


def _sharded_host_worker(rank, world, port, out_dir):
    """The N > 1 host flow on CPU: shared host buffers, the exchange (numpy restatement), per-slab compute
    (oracle), normal records to their home range, contours per rank, ONE result assembled on rank 0."""
    import torch.distributed as dist
    from exchange_model import exchange_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, k, halo, S = 24000, 16, 12.0, 10
    starts = parallel.index_ranges(n, world)
    src = parallel.SharedHost(dist, rank, world, n * 32, "t_src_%d" % port)
    cloud_h = src.array(np.float32, (n, 8))
    if rank == 0:
        cloud_h[...] = synth.panel(n, 5)
    dist.barrier()
    lay = parallel.host_region_layout(world, n, 16, S, 8000)
    dst = parallel.SharedHost(dist, rank, world, lay["total_bytes"], "t_dst_%d" % port)
    normals_h = dst.array(np.float32, (n, 4))
    # exchange: every rank evaluates the restatement on the shared cloud and keeps its own slab
    slabs, info = exchange_model([cloud_h[starts[r]:starts[r + 1]] for r in range(world)], starts, halo)
    xyz, w = slabs[rank]
    oc = po.OracleCloud(np.ascontiguousarray(xyz))
    nrm, _ = oc.normals(k=k)
    own = w >= 0
    # "home" delivery: this rank's records land in other ranks' index ranges of the one result array
    normals_h[w[own]] = nrm[own]
    planes = synth.even_planes(cloud_h, S)
    pos = parallel.owned_planes(planes, info["cuts"], rank)
    off, y, x, z = oc.slice_contours(planes[pos], "B")
    at = lay["normals_bytes"] + rank * lay["region_bytes"]
    dst.array(np.int64, (len(pos) + 1,), at)[...] = off
    for j, a in enumerate((y, x, z)):
        dst.array(np.float64, (len(a),), at + lay["off_bytes"] + j * lay["arr_bytes"])[...] = a
    all_pos = [None] * world
    dist.all_gather_object(all_pos, pos.tolist())
    dist.barrier()
    if rank == 0:
        per_rank = []
        for r in range(world):
            at = lay["normals_bytes"] + r * lay["region_bytes"]
            o = dst.array(np.int64, (len(all_pos[r]) + 1,), at).copy()
            arrs = [dst.array(np.float64, (int(o[-1]),), at + lay["off_bytes"] + j * lay["arr_bytes"]).copy() for j in range(3)]
            per_rank.append((np.asarray(all_pos[r], np.int64), o) + tuple(arrs))
        goff, gy, gx, gz = parallel.assemble_contours(S, per_rank)
        np.savez(os.path.join(out_dir, "one_host.npz"), normals=normals_h.copy(), off=goff, y=gy, z=gz)
    dist.barrier()
    del cloud_h, normals_h
    src.close()
    dst.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("memfd", [False, True])
def test_gloo_sharded_flow_delivers_one_host_result(tmp_path, memfd, monkeypatch):
    """memfd = True: the backing object is an anonymous memfd reached through /proc/<pid>/fd (what SharedHost
    falls back to when /dev/shm is too small for the cloud)."""
    import torch.multiprocessing as mp
    if memfd:
        monkeypatch.setenv("PPP_SHM_MEMFD", "1")
    port = 31500 + (os.getpid() % 2000) + (7 if memfd else 0)
    mp.spawn(_sharded_host_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(str(tmp_path / "one_host.npz"))
    cloud = synth.panel(24000, 5)
    oc = po.OracleCloud(cloud)
    ref_n, _ = oc.normals(k=16)
    ref_c = oc.slice_contours(synth.even_planes(cloud, 10), "B")
    assert np.array_equal(got["normals"].view(np.uint32), ref_n.view(np.uint32))
    assert np.array_equal(got["off"], ref_c[0]) and np.array_equal(got["y"], ref_c[1]) and np.array_equal(got["z"], ref_c[3])
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("ppp_t_")]      # backing objects removed


def test_steffen_spline_restatements_agree():
    """Spline.h's gsl_interp_steffen: the product's numpy mirror equals the oracle restatement bit for
    bit, interpolates the nodes, stays monotone between monotone nodes (Steffen 1990) and rejects what
    GSL aborts on."""
    from polishpathplanning_b200 import reference_api as ra
    rng = np.random.default_rng(0)
    y = np.cumsum(rng.random(40) + 0.05)
    x = np.full(40, 12.5)
    z = np.sin(y / 3) * 5 + rng.normal(0, 0.05, 40)
    sp = ra.Spline(y, x, z)
    q = np.linspace(y[0], y[-1], 1001)
    p = sp.point(q)
    assert np.array_equal(p[:, 2].view(np.uint64), po.steffen_eval(y, z, q).view(np.uint64))
    assert np.array_equal(p[:, 0], np.full(1001, 12.5)) and np.array_equal(p[:, 1], q)
    assert np.allclose(sp.point(y)[:, 2], z, rtol=0, atol=1e-12)
    zm = np.cumsum(rng.random(40))                       # monotone data -> monotone spline
    pm = ra.Spline(y, x, zm).point(q)[:, 2]
    assert np.all(np.diff(pm) >= -1e-12)
    assert np.isnan(sp.point([y[0] - 1.0, y[-1] + 1.0])[:, 2]).all()
    with pytest.raises(ValueError):
        ra.Spline(y[:2], x[:2], z[:2])
    with pytest.raises(ValueError):
        ra.Spline(y[::-1], x, z)
    with pytest.raises(ValueError):
        po.steffen_eval(y[:2], z[:2], q)


def test_sor_threshold_pass_restatements_agree():
    """The host mirror of StatisticalOutlierRemoval's sequential statistics equals the oracle's loop."""
    from polishpathplanning_b200 import reference_api as ra
    rng = np.random.default_rng(11)
    for n in (2, 7, 1000, 250000):
        d = rng.gamma(9.0, 0.3, n).astype(np.float32)
        d[rng.random(n) < 0.01] = 0.0
        nv = int((d != 0).sum()) or 1
        for mul in (1.0, 0.25, 3.0):
            for neg in (False, True):
                kept, thr = ra.sor_select(d, nv, mul, neg)
                okept, othr = po.OracleCloud.sor_select(d, nv, mul, neg)
                assert np.array_equal(kept, okept)
                assert thr == othr or (np.isnan(thr) and np.isnan(othr))


def test_host_region_layout_sections_are_disjoint_and_aligned():
    """The one host result buffer: normals, then one contour region per rank."""
    for world, n_total, stride, S_cap, node_cap in ((2, 2_000_000, 32, 142, 270_000), (8, 8_000_001, 16, 71, 65_536), (3, 5, 32, 0, 0)):
        lay = parallel.host_region_layout(world, n_total, stride, S_cap, node_cap)
        spans = [(0, n_total * stride)]
        for r in range(world):
            at = lay["normals_bytes"] + r * lay["region_bytes"]
            spans += [(at, at + (S_cap + 1) * 8)]
            spans += [(at + lay["off_bytes"] + j * lay["arr_bytes"], at + lay["off_bytes"] + j * lay["arr_bytes"] + node_cap * 8)
                      for j in range(3)]
        assert all(a % 256 == 0 for a, _ in spans)
        spans.sort()
        assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
        assert spans[-1][1] <= lay["total_bytes"]


class _FakeExchange:
    """Stands in for parallel.Exchange in the rank-protocol test: records the calls; `fail` makes the
    creation or the connection raise on this rank."""
    fail = None
    log = []

    def __init__(self, ctx, rank, world, starts, cap_recv, S_cap, node_cap, normal_stride_bytes):
        if _FakeExchange.fail == "create":
            raise RuntimeError("no memory")
        self.rank, self.world, self.caps, self.starts = rank, world, (cap_recv, S_cap, node_cap), list(starts)
        _FakeExchange.log.append(("create", rank))

    def handle(self):
        return bytes([self.rank]) * 64

    def connect_ipc(self, handles):
        if _FakeExchange.fail == "connect":
            raise RuntimeError("cannot map")
        assert [h[0] for h in handles] == list(range(self.world)) and all(len(h) == 64 for h in handles)
        _FakeExchange.log.append(("connect", self.rank))

    def close(self, dist=None):
        if dist is not None:
            dist.barrier()
        _FakeExchange.log.append(("close", self.rank))


def _exchange_protocol_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cpu = torch.device("cpu")
    # 1. success: one set of capacities on all ranks (maxima of the requests), handles in rank order
    _FakeExchange.fail, _FakeExchange.log = None, []
    ex = parallel.Exchange.over_dist(None, dist, cpu, rank, world, 1001, cap_recv=700 + 10 * rank, S_cap=7 - rank,
                                     node_cap=100 + 50 * rank, factory=_FakeExchange)
    assert ex.caps == (700 + 10 * (world - 1), 7, 100 + 50 * (world - 1))
    assert ex.starts == [0, 500, 1001]
    assert _FakeExchange.log == [("create", rank), ("connect", rank)]
    # 2. a failure anywhere is raised on EVERY rank, nothing left open
    for fail_rank, kind in ((0, "create"), (1, "connect")):
        _FakeExchange.fail, _FakeExchange.log = (kind if rank == fail_rank else None), []
        with pytest.raises(RuntimeError):
            parallel.Exchange.over_dist(None, dist, cpu, rank, world, 1001, 700, 7, 100, factory=_FakeExchange)
        created = [c for c in _FakeExchange.log if c[0] == "create"]
        closed = [c for c in _FakeExchange.log if c[0] == "close"]
        assert len(created) == len(closed)
    with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
        f.write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_exchange_setup_protocol(tmp_path):
    """Exchange.over_dist on CPU (gloo, world 2) with a stand-in exchange: common capacities, index ranges,
    handle all-gather in rank order, and collective failure when any rank cannot create / connect."""
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_exchange_protocol_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(str(tmp_path / "ok0")) and os.path.exists(str(tmp_path / "ok1"))
