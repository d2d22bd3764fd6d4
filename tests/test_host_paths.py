"""CPU tests of the per-path host logic: the BATCHED Area2Cloud / compute_boundary / bisection /
dynamic_adjust_path and getPath way-point code of polishpathplanning_b200/reference_api.py against a node-by-node
transcription of the reference's loops (tests/ref_scalar.py), both running on the oracle-backed device stand-in:
identical results, and one device call where the reference makes one per node."""
import numpy as np
import pytest

from oracle_dev import OracleDev
import ref_scalar as rs
from polishpathplanning_b200 import reference_api as ra, synth


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.fixture(scope="module")
def gen2():
    cloud_m = synth.to_pointxyzrgb(synth.panel_metres(30000, 12))
    pg = ra.path_generater(cloud_m, 15.0, backend=OracleDev)
    pg.estimate_normal()
    pg.Contact_Path_Generation(adjust=False)
    return pg


def test_area2cloud_batched_equals_node_by_node(gen2):
    pg = gen2
    path = ra.Spline(*pg.Path_set[2])
    nodes = path.point(np.linspace(path.miny() + 2, path.bigy() - 2, 37))
    nodes[5, 0] += 400.0                          # a node far off the workpiece
    dev = pg._dev()
    for key in (False, True):
        pg.coverage_flag = np.zeros(pg.cloud.shape[0], np.uint8)
        dev.calls.clear()
        got = pg.Area2Cloud(nodes, False, key)
        assert dev.calls["principal_curvatures"] == 1          # one device call for all 37 nodes
        flags = np.zeros(pg.cloud.shape[0], np.uint8)
        want = np.stack([rs.area2cloud(dev, pg.cloud_with_normals, flags, p, key, pg.toolRadius) for p in nodes])
        assert np.array_equal(_bits(got), _bits(want))
        assert np.array_equal(pg.coverage_flag, flags)
        if not key:
            assert 0 < flags.sum() < flags.shape[0]
    # "up" boundary lies at larger x than the "down" one
    up, down = pg.Area2Cloud(nodes[:5], False, False), pg.Area2Cloud(nodes[:5], False, True)
    assert np.all(up[:, 0] > down[:, 0])


def test_compute_boundary_and_bisection_equal_node_by_node(gen2):
    pg = gen2
    dev = pg._dev()
    pre, cur = ra.Spline(*pg.Path_set[2]), ra.Spline(*pg.Path_set[3])
    pg.coverage_flag = np.zeros(pg.cloud.shape[0], np.uint8)
    dev.calls.clear()
    b = pg.compute_boundary(pre)
    assert dev.calls["principal_curvatures"] == 1
    flags = np.zeros(pg.cloud.shape[0], np.uint8)
    dev.calls.clear()
    bs = rs.compute_boundary(dev, pg.cloud_with_normals, flags, pre, pg.toolRadius, ra.Spline)
    assert dev.calls["principal_curvatures"] > 20               # the reference's way: one query per node
    assert b is not None and bs is not None
    assert np.array_equal(b.y, bs.y) and all(np.array_equal(u, v) for u, v in zip(b._cx + b._cz, bs._cx + bs._cz))
    assert np.array_equal(pg.coverage_flag, flags)
    assert b.miny() < pre.miny() and b.bigy() > pre.bigy() - 20      # extended by 20 at both ends
    # nodes of a path TWO planes further on: their "down" boundary is far from this boundary, so they must move
    far = ra.Spline(*pg.Path_set[4])
    nodes = far.point(np.linspace(far.miny() + 6, far.bigy() - 6, 23))
    dev.calls.clear()
    got = pg.bisection(nodes, b)
    assert dev.calls["principal_curvatures"] <= 6                # at most six rounds, whatever the node count
    want = np.stack([rs.bisection(dev, pg.cloud_with_normals, flags, p, bs, pg.toolRadius) for p in nodes])
    assert np.array_equal(_bits(got), _bits(want))
    assert np.array_equal(got[:, 1:], nodes[:, 1:])             # only x moves
    assert (got[:, 0] != nodes[:, 0]).any()


def test_dynamic_adjust_path_equals_node_by_node(gen2):
    pg = gen2
    dev = pg._dev()
    pre, cur = ra.Spline(*pg.Path_set[1]), ra.Spline(*pg.Path_set[2])
    pg.coverage_flag = np.zeros(pg.cloud.shape[0], np.uint8)
    got = pg.dynamic_adjust_path(cur, pre)
    flags = np.zeros(pg.cloud.shape[0], np.uint8)
    want = rs.dynamic_adjust_path(dev, pg.cloud, pg.cloud_with_normals, flags, cur, pre, pg.toolRadius, ra.Spline)
    assert np.array_equal(got.y, want.y) and all(np.array_equal(u, v) for u, v in zip(got._cx + got._cz, want._cx + want._cz))
    assert np.array_equal(pg.coverage_flag, flags)
    # the new nodes are cloud points
    assert np.isin(got.y, pg.cloud[:, 1].astype(np.float64)).all()


def test_contact_path_generation_with_adjustment_runs_the_reference_sweep():
    cloud_m = synth.to_pointxyzrgb(synth.panel_metres(12000, 3))
    pg = ra.path_generater(cloud_m, 15.0, backend=OracleDev)
    pg.estimate_normal()
    planes = pg.Contact_Path_Generation(adjust=True)
    assert len(pg.Path_splines) == len(planes) >= 3
    raw0 = ra.Spline(*pg.Path_set[0])
    assert np.array_equal(pg.Path_splines[0].y, raw0.y)          # the first path is never adjusted
    assert any(not np.array_equal(pg.Path_splines[s].y, pg.Path_set[s][0]) for s in range(1, len(planes)))
    assert 0.05 < pg.get_coverage() <= 1.0


def test_getpath_waypoints_equal_waypoint_by_waypoint():
    cloud_m = synth.to_pointxyzrgb(synth.panel_metres(20000, 7))
    sp = ra.SectPath(cloud_m, 12.0, backend=OracleDev)
    sp.GenPath()
    sp.estimate_normal()
    th = np.float32(0.3)
    T = np.eye(4, dtype=np.float32)
    T[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
    T[:3, 3] = [5.0, -3.0, 1.5]
    for Tr in (None, T):
        dev = sp._dev()
        dev.calls.clear()
        xyz, idx, rot, tails = sp.getPath_waypoints(7.5, Tr)
        assert dev.calls["knn"] == 1
        inv = np.linalg.inv(Tr).astype(np.float32) if Tr is not None else np.eye(4, dtype=np.float32)
        wx, wi, wr, wt = rs.getpath_waypoints(dev, sp.cloud_with_normals, sp.splines(), 7.5, inv)
        assert np.array_equal(_bits(xyz), _bits(wx)) and np.array_equal(idx, wi) and np.array_equal(tails, wt)
        assert np.array_equal(_bits(rot), _bits(wr))
        assert len(tails) == len(sp.splines()) - 2
    # boustrophedon: consecutive paths run in opposite y directions
    xyz, _, _, tails = sp.getPath_waypoints(7.5)
    a, b = xyz[:tails[0] + 1, 1], xyz[tails[0] + 1:tails[1] + 1, 1]
    assert np.all(np.diff(a) > 0) and np.all(np.diff(b) < 0)
