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


def _redistribute_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cloud = synth.panel(12000, 17)
    cloud[100, 1] = np.nan
    n = cloud.shape[0]
    a, b = (n * rank) // world, (n * (rank + 1)) // world
    local, g, owned, cuts = parallel.redistribute(dist, torch.from_numpy(cloud[a:b].copy()), a, rank, world, 12.0)
    local, g, owned = local.numpy(), g.numpy(), owned.numpy()
    assert np.all(np.diff(g) > 0)                                    # ascending global index
    assert np.array_equal(local[:, :5].view(np.uint32), cloud[g][:, :5].view(np.uint32))   # records intact
    oc = po.OracleCloud(local)
    nrm, _ = oc.normals(k=16)
    _, d2 = oc.knn(16)
    assert len(parallel.halo_violations(local, owned, d2[:, -1], cuts, rank, 12.0)) == 0
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), g=g[owned], nrm=nrm[owned])
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_redistribute_exchange(tmp_path):
    """The all-to-all-v halo exchange (CPU tensors over gloo): ownership partitions the cloud and the
    sharded normals equal the single-process ones bit for bit."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_redistribute_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cloud = synth.panel(12000, 17)
    cloud[100, 1] = np.nan
    ref, _ = po.OracleCloud(cloud).normals(k=16)
    parts = [np.load(str(tmp_path / ("r%d.npz" % r))) for r in range(2)]
    allg = np.concatenate([p["g"] for p in parts])
    assert np.array_equal(np.sort(allg), np.arange(cloud.shape[0]))  # every point owned exactly once
    full = parallel.assemble_normals(cloud.shape[0], 4, [(p["g"], p["nrm"]) for p in parts])
    assert np.array_equal(full.view(np.uint32), ref.view(np.uint32))


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


def test_peer_sink_layout_sections_are_disjoint_and_aligned():
    """The one buffer every rank's kernels store into: normals, per-rank node regions, flags."""
    for world, n_total, node_cap, S_cap in ((2, 2_000_000, 270_000, 142), (8, 8_000_001, 65_536, 71), (3, 5, 1, 0)):
        lay = parallel.peer_sink_layout(world, n_total, node_cap, S_cap)
        spans = [(0, n_total * 16)]
        for r in range(world):
            at = lay["normals_bytes"] + r * lay["region_bytes"]
            spans += [(at, at + (S_cap + 1) * 8)]
            spans += [(at + lay["off_bytes"] + j * lay["arr_bytes"], at + lay["off_bytes"] + j * lay["arr_bytes"] + node_cap * 8)
                      for j in range(3)]
        spans += [(lay["flags_at"] + 128 * r, lay["flags_at"] + 128 * r + 4) for r in range(world)]
        assert all(a % 256 == 0 or a >= lay["flags_at"] for a, _ in spans)
        assert all(a % 128 == 0 for a, _ in spans)
        spans.sort()
        assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
        assert spans[-1][1] <= lay["total_bytes"]


class _FakePeerCtx:
    """Stands in for api.Context in the PeerSink protocol test: records the calls, hands out fake
    addresses; `fail` makes the owner's allocation (or a peer's mapping) raise."""

    def __init__(self, rank, fail=None):
        self.rank, self.fail, self.calls = rank, fail, []

    def peer_buffer_alloc(self, nbytes):
        if self.fail == "alloc":
            raise RuntimeError("no IPC")
        self.calls.append(("alloc", nbytes))
        return 0x10000000, bytes(range(64))

    def peer_buffer_open(self, handle):
        if self.fail == "open":
            raise RuntimeError("cannot map")
        assert handle == bytes(range(64))
        self.calls.append(("open",))
        return 0x20000000

    def peer_buffer_close(self, p):
        self.calls.append(("close", p))

    def peer_buffer_free(self, p):
        self.calls.append(("free", p))

    def signal(self, p, v):
        self.calls.append(("signal", p, v))

    def wait(self, p, v):
        self.calls.append(("wait", p, v))


def _peer_sink_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cpu = torch.device("cpu")
    # 1. success: one layout on all ranks (capacities are the maxima of the requests), flags, teardown order
    ctx = _FakePeerCtx(rank)
    sink = parallel.PeerSink(ctx, dist, cpu, rank, world, n_total=1000, node_cap=100 + 50 * rank, S_cap=7 - rank)
    assert (sink.n_total, sink.node_cap, sink.S_cap) == (1000, 100 + 50 * (world - 1), 7)
    lay = parallel.peer_sink_layout(world, 1000, sink.node_cap, 7)
    assert sink.flag(1) - sink.base == lay["flags_at"] + 128
    assert sink.region(1)["y"] - sink.base == lay["normals_bytes"] + lay["region_bytes"] + lay["off_bytes"]
    sink.delivered(3)
    want = [("signal", sink.flag(rank), 3)] + ([("wait", sink.flag(r), 3) for r in range(1, world)] if rank == 0 else [])
    assert ctx.calls[-len(want):] == want
    base = sink.base
    sink.close()
    assert ctx.calls[-1] == (("free", base) if rank == 0 else ("close", base))
    # 2. a failure anywhere is raised on EVERY rank (so callers can fall back together), nothing left mapped
    for fail_rank, kind in ((0, "alloc"), (1, "open")):
        ctx = _FakePeerCtx(rank, fail=kind if rank == fail_rank else None)
        try:
            parallel.PeerSink(ctx, dist, cpu, rank, world, 1000, 100, 7)
            raised = False
        except RuntimeError:
            raised = True
        assert raised
        opened = [c for c in ctx.calls if c[0] in ("alloc", "open")]
        released = [c for c in ctx.calls if c[0] in ("free", "close")]
        assert len(opened) == len(released)
    with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
        f.write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_peer_sink_protocol(tmp_path):
    """PeerSink's rank protocol on CPU (gloo, world 2) with a fake context: common layout, flag
    signalling, teardown, and collective failure when a mapping is unavailable."""
    import torch.multiprocessing as mp
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_peer_sink_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(str(tmp_path / "ok0")) and os.path.exists(str(tmp_path / "ok1"))
