"""The C++ mirror of the reference classes (polishpathplanning_b200/host) run as the reference's
own ./main and config.txt flows, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import ppp_oracle as po
from polishpathplanning_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "polishpathplanning_b200", "host")
EXE = os.path.join(HOST, "ppp_main")


def _build():
    subprocess.check_call(["make", "-s", "-C", HOST])
    assert os.path.exists(EXE)


def _load(prefix):
    off = np.fromfile(prefix + ".off.i64", np.int64)
    return (off, np.fromfile(prefix + ".y.f64", np.float64), np.fromfile(prefix + ".x.f64", np.float64),
            np.fromfile(prefix + ".z.f64", np.float64), np.fromfile(prefix + ".normals.f32", np.float32).reshape(-1, 8))


def _check_normals(g, o):
    gn = np.concatenate([g[:, 0:3], g[:, 4:5]], axis=1)
    assert np.array_equal(np.isnan(gn[:, 0]), np.isnan(o[:, 0]))
    ok = ~np.isnan(o[:, 0])
    assert np.abs(gn[ok] - o[ok]).max() <= 1e-5


def test_main_flow(tmp_path):
    """./main workpiece.pcd: ctor scaling, estimate_normal (r = 2.5), Contact_Path_Generation sweep (R = 15)."""
    _build()
    pcd = str(tmp_path / "workpiece.pcd")
    synth.write_pcd(pcd, synth.to_pointxyzrgb(synth.panel_metres(100000, 0)))   # BASELINE configs[0] size
    out = str(tmp_path / "main")
    r = subprocess.run([EXE, pcd, out], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Number of paths" in r.stdout
    off, y, x, z, nrm = _load(out)
    cloud = synth.panel(100000, 0)
    oc = po.OracleCloud(cloud)
    _check_normals(nrm, oc.normals(radius=2.5)[0])
    mn, mx = oc.minmax()
    planes = po.planes("gen2_contact", mn[0], mx[0], 15.0)
    ooff, oy, ox, oz = oc.slice_contours(planes, "A")
    assert np.array_equal(off, ooff) and np.array_equal(y, oy) and np.array_equal(x, ox) and np.array_equal(z, oz)
    assert os.path.exists(str(tmp_path / "output.csv"))   # the reference appends its sweep time there


def test_sectpath_config_flow(tmp_path):
    """connect-style flow: SectPath(config.txt, cloud) with the shipped config values (Tool_Radius 12)."""
    _build()
    pcd = str(tmp_path / "workpiece.pcd")
    synth.write_pcd(pcd, synth.to_pointxyzrgb(synth.panel_metres(60000, 3)), binary=False)
    cfg = str(tmp_path / "config.txt")
    with open(cfg, "w") as f:
        f.write("Tool_Radius = 12\npathFile = path.txt\n# comment\nPathResolution = 5\nSmooth = false\n"
                "Alignment = false\nChangeRange = true\nRemoveOutlier = false\nDynamic_adjustment = false\n")
    out = str(tmp_path / "sect")
    r = subprocess.run([EXE, "--sect", cfg, pcd, out], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0, r.stderr
    off, y, x, z, nrm = _load(out)
    cloud = synth.panel(60000, 3)
    oc = po.OracleCloud(cloud)
    _check_normals(nrm, oc.normals(radius=2.5)[0])
    mn, mx = oc.minmax()
    planes = po.planes("sectpath", mn[0], mx[0], 12.0)
    ooff, oy, ox, oz = oc.slice_contours(planes, "B")
    assert np.array_equal(off, ooff) and np.array_equal(y, oy) and np.array_equal(z, oz)


def test_missing_file_does_not_abort(tmp_path):
    _build()
    r = subprocess.run([EXE, str(tmp_path / "nope.pcd")], capture_output=True, text=True, cwd=str(tmp_path), timeout=120)
    assert r.returncode == 0 and "read file" in r.stderr


def test_slicing_method_flow(tmp_path):
    """path_generater::slicing_method (src/Path_Generation.cpp:282-321): first plane min.x + int(2R)/2."""
    _build()
    pcd = str(tmp_path / "workpiece.pcd")
    synth.write_pcd(pcd, synth.to_pointxyzrgb(synth.panel_metres(50000, 9)))
    out = str(tmp_path / "sl")
    r = subprocess.run([EXE, "--slicing", pcd, out], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0, r.stderr
    assert "number of paths" in r.stdout
    off = np.fromfile(out + ".off.i64", np.int64)
    y = np.fromfile(out + ".y.f64", np.float64)
    z = np.fromfile(out + ".z.f64", np.float64)
    oc = po.OracleCloud(synth.panel(50000, 9))
    mn, mx = oc.minmax()
    planes = po.planes("gen2_slicing", mn[0], mx[0], 15.0)
    ooff, oy, ox, oz = oc.slice_contours(planes, "A")
    assert np.array_equal(off, ooff) and np.array_equal(y, oy) and np.array_equal(z, oz)


def test_sectpath_remove_outlier_flow(tmp_path):
    """config.txt with RemoveOutlier = true: SOR(50, 1.0) in the constructor, then normals + sweep."""
    _build()
    pcd = str(tmp_path / "workpiece.pcd")
    metres = synth.to_pointxyzrgb(synth.panel_metres(40000, 9))
    metres[::400, 2] += np.float32(0.03)                      # 30 mm outliers
    synth.write_pcd(pcd, metres)
    cfg = str(tmp_path / "config.txt")
    with open(cfg, "w") as f:
        f.write("Tool_Radius = 12\nChangeRange = true\nRemoveOutlier = true\n")
    out = str(tmp_path / "sect")
    r = subprocess.run([EXE, "--sect", cfg, pcd, out], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0, r.stderr
    off, y, x, z, nrm = _load(out)
    cloud = metres.copy()
    cloud[:, :3] = metres[:, :3] * np.float32(1000)
    keep, _ = po.OracleCloud.sor_select(*po.OracleCloud(cloud).sor_mean_distances(50), 1.0)
    assert 0 not in keep and keep.shape[0] < cloud.shape[0]
    oc = po.OracleCloud(np.ascontiguousarray(cloud[keep]))
    assert nrm.shape[0] == keep.shape[0]
    _check_normals(nrm, oc.normals(radius=2.5)[0])
    mn, mx = oc.minmax()
    planes = po.planes("sectpath", mn[0], mx[0], 12.0)
    ooff, oy, ox, oz = oc.slice_contours(planes, "B")
    assert np.array_equal(off, ooff) and np.array_equal(y, oy) and np.array_equal(z, oz)
