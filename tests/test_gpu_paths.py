"""GPU tests of the per-path consumers through the reference-shaped Python interface: the
device-backed batched Area2Cloud / compute_boundary / bisection / dynamic_adjust_path / getPath way-points against
the same host logic running on the oracle stand-in."""
import numpy as np
import pytest

from oracle_dev import OracleDev
from polishpathplanning_b200 import reference_api as ra, synth

pytestmark = pytest.mark.gpu


def test_gen2_adjusted_sweep_matches_oracle_backend(ctx):
    cloud_m = synth.to_pointxyzrgb(synth.panel_metres(30000, 12))
    g = ra.path_generater(cloud_m, 15.0, ctx=ctx)
    o = ra.path_generater(cloud_m, 15.0, backend=OracleDev)
    for pg in (g, o):
        pg.estimate_normal()
        pg.Contact_Path_Generation(adjust=False)
    assert len(g.Path_set) == len(o.Path_set)
    for a, b in zip(g.Path_set, o.Path_set):
        assert all(np.array_equal(u, v) for u, v in zip(a, b))
    path = ra.Spline(*g.Path_set[2])
    nodes = path.point(np.linspace(path.miny() + 2, path.bigy() - 2, 64))
    for key in (False, True):
        for pg in (g, o):
            pg.coverage_flag = np.zeros(pg.cloud.shape[0], np.uint8)
        bg, bo = g.Area2Cloud(nodes, False, key), o.Area2Cloud(nodes, False, key)
        # principal curvatures go through float32 eigen arithmetic on both sides; the ellipse axes amplify their
        # last-bit differences a little (sqrt(2 r d)): tolerance, not bits
        assert np.abs(bg - bo).max() < 2e-3
        if not key:
            assert (g.coverage_flag != o.coverage_flag).mean() < 2e-3 and g.coverage_flag.sum() > 0
    pre, far = ra.Spline(*g.Path_set[2]), ra.Spline(*g.Path_set[4])
    Bg, Bo = g.compute_boundary(pre), o.compute_boundary(pre)
    assert Bg is not None and Bo is not None and len(Bg.y) == len(Bo.y) and np.abs(Bg.y - Bo.y).max() < 2e-3
    q = far.point(np.linspace(far.miny() + 6, far.bigy() - 6, 40))
    ng, no = g.bisection(q, Bg), o.bisection(q, Bo)
    assert np.abs(ng - no).max() < 5e-3 and (ng[:, 0] != q[:, 0]).any()
    ag, ao = g.dynamic_adjust_path(far, pre), o.dynamic_adjust_path(far, pre)
    # snapped to cloud points: the same points except where a node sits within the tolerance of a Voronoi border
    same = np.isin(ag.y, ao.y).mean()
    assert same > 0.9
    g._invalidate()


def test_sectpath_waypoints_match_oracle_backend(ctx):
    cloud_m = synth.to_pointxyzrgb(synth.panel_metres(20000, 7))
    g = ra.SectPath(cloud_m, 12.0, ctx=ctx)
    o = ra.SectPath(cloud_m, 12.0, backend=OracleDev)
    for sp in (g, o):
        sp.GenPath()
        sp.estimate_normal()
    xg, ig, rg, tg = g.getPath_waypoints(7.5)
    xo, io, ro, to = o.getPath_waypoints(7.5)
    assert np.array_equal(xg, xo) and np.array_equal(ig, io) and np.array_equal(tg, to)   # same splines, same snaps
    assert np.abs(rg - ro).max() <= 2e-5                                                 # frames from normals (<= 1e-5 each)
    assert xg.shape[0] > 30
    g._invalidate()
