"""Parity on workpieces that are NOT height fields (a closed cylinder, a box with vertical walls): the grid
index spans two axes, so whole strips of a vertical wall share a cell column -- results must stay exact."""
import numpy as np
import pytest

from oracle import ppp_oracle as po
from polishpathplanning_b200 import api, synth

pytestmark = pytest.mark.gpu

SHAPES = {"cylinder": lambda n, s: synth.cylinder(n, s), "box": lambda n, s: synth.box_with_walls(n, s),
          "cylinder_dense": lambda n, s: synth.cylinder(n, s, density=6.0)}


def _normals_close(g, o):
    gn = np.concatenate([g[:, 0:3], g[:, 4:5]], axis=1)
    nan_g, nan_o = np.isnan(gn[:, 0]), np.isnan(o[:, 0])
    assert np.array_equal(nan_g, nan_o)
    ok = ~nan_o
    # sign alignment is part of the result (flipNormalTowardsViewpoint), so compare as is
    assert np.abs(gn[ok] - o[ok]).max() <= 1e-5


@pytest.mark.parametrize("shape,n", [("cylinder", 100000), ("box", 100000), ("cylinder_dense", 60000)])
def test_closed_and_steep_workpieces_match_oracle(ctx, shape, n):
    cloud = SHAPES[shape](n, 3)
    gc = api.Cloud(ctx, cloud)
    oc = po.OracleCloud(cloud)
    for k in (16, 32):
        g, gi = gc.normals_knn(k, return_idx=True)
        oi, _ = oc.knn(k, want_d2=False)
        assert np.array_equal(gi, oi), "%s: kNN(k=%d) index lists differ" % (shape, k)
        o, _ = oc.normals(k=k)
        _normals_close(g, o)
    g = gc.normals_radius(2.5)
    o, _ = oc.normals(radius=2.5)
    _normals_close(g, o)
    mn, mx = gc.bbox()
    omn, omx = oc.minmax()
    assert np.array_equal(mn, omn) and np.array_equal(mx, omx)
    planes = po.planes("sectpath", mn[0], mx[0], 9.0)
    goff, gidx = gc.slice_bands(planes)
    ooff, oidx = oc.slice_bands(planes)
    assert np.array_equal(goff, ooff) and np.array_equal(gidx, oidx)
    for mode in ("A", "B"):
        goff, gy, gx, gz = gc.slice_contours(planes, mode)
        ooff, oy, ox, oz = oc.slice_contours(planes, mode)
        assert np.array_equal(goff, ooff), "%s: contour node counts differ (%s)" % (shape, mode)
        assert np.array_equal(gy.view(np.uint64), oy.view(np.uint64)) and np.array_equal(gz.view(np.uint64), oz.view(np.uint64))
    gc.close()
