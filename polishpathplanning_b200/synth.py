"""Seeded synthetic workpiece clouds and a minimal PCD reader/writer.

The reference ships no data (its /PCD directory is git-ignored, /root/reference/.gitignore:3), so
benchmarks and parity tests use the deterministic "panel" of SURVEY.md §8(d): a freeform height
field sampled at ~1 point/mm^2, stored in METRES like the reference's input files, then scaled by
the reference's own constructor loop (x,y,z *= 1000 in float32, rgb = 255;
/root/reference/src/Path_Generation.cpp:23-31) so that loop defines the bits every consumer sees.

Layout: pcl::PointXYZRGB = 32 bytes = float x,y,z,pad(1.0f); uint32 rgba; 3 x pad.  Clouds are
handed around as float32 arrays of shape (N, 8) (column 4 holds the rgba bits).
"""
import numpy as np

POINT_STRIDE_FLOATS = 8  # pcl::PointXYZRGB, 32 B
NORMAL_STRIDE_FLOATS = 8  # pcl::Normal, 32 B: nx,ny,nz,pad, curvature,pad,pad,pad


def panel_metres(n, seed=0, density=1.0, noise_sigma=0.0):
    """(n, 3) float32 in metres. side L = sqrt(n/density) mm; uniform (x, y); curved z."""
    rng = np.random.default_rng(seed)
    L = np.sqrt(n / density)
    x = rng.random(n) * L
    y = rng.random(n) * L
    z = 20.0 * np.sin(2 * np.pi * x / 400.0) * np.cos(2 * np.pi * y / 300.0) + 1e-4 * (x - L / 2) ** 2
    if noise_sigma > 0:
        z = z + rng.normal(0.0, noise_sigma, n)
    xyz_mm = np.stack([x, y, z], axis=1).astype(np.float32)
    return (xyz_mm / np.float32(1000.0)).astype(np.float32)


def to_pointxyzrgb(xyz, rgb=0):
    """(N,3) float32 -> (N,8) float32 PointXYZRGB records."""
    n = xyz.shape[0]
    out = np.zeros((n, POINT_STRIDE_FLOATS), np.float32)
    out[:, 0:3] = xyz
    out[:, 3] = 1.0
    out[:, 4] = np.full(n, rgb, np.uint32).view(np.float32)
    return out


def reference_ctor_scale(cloud):
    """In place: x,y,z *= 1000 (float32), r=g=b=255  (src/Path_Generation.cpp:23-31)."""
    cloud[:, 0:3] *= np.float32(1000.0)
    rgba = cloud[:, 4].view(np.uint32)
    rgba[:] = (rgba & np.uint32(0xFF000000)) | np.uint32(0x00FFFFFF)
    return cloud


def panel(n, seed=0, density=1.0, noise_sigma=0.0):
    """The cloud exactly as the reference's constructor leaves it: (n, 8) float32, millimetres."""
    return reference_ctor_scale(to_pointxyzrgb(panel_metres(n, seed, density, noise_sigma)))


def _records_mm(xyz_mm):
    """(n, 3) millimetre coordinates -> the cloud as the reference's constructor leaves it (stored in metres
    as float32, then x, y, z *= 1000 in float32)."""
    xyz_m = (np.asarray(xyz_mm, np.float64).astype(np.float32) / np.float32(1000.0)).astype(np.float32)
    return reference_ctor_scale(to_pointxyzrgb(xyz_m))


def cylinder(n, seed=0, density=1.0, aspect=4.0):
    """A CLOSED workpiece: the lateral surface of a cylinder whose axis is y, ~density points/mm^2, length =
    aspect x radius.  Not a height field: above every (x, y) there are two sheets, and near x = +-R the
    surface is vertical -- a grid over two axes piles whole strips of it into single columns."""
    rng = np.random.default_rng(seed)
    R = np.sqrt(n / density / (2 * np.pi * aspect))
    L = aspect * R
    th = rng.random(n) * 2 * np.pi
    y = rng.random(n) * L
    x = R * np.cos(th) + R          # x in [0, 2R]: planes x = const cut it into two arcs
    z = R * np.sin(th)
    return _records_mm(np.stack([x, y, z], axis=1))


def box_with_walls(n, seed=0, density=1.0, wall=0.35):
    """An open box: a flat floor of side L with four VERTICAL walls of height wall x L (a steep-sided
    workpiece).  Points are spread by area at ~density points/mm^2."""
    rng = np.random.default_rng(seed)
    L = np.sqrt(n / density / (1 + 4 * wall))
    H = wall * L
    n_floor = int(round(n / (1 + 4 * wall)))
    n_wall = n - n_floor
    fx, fy = rng.random(n_floor) * L, rng.random(n_floor) * L
    floor = np.stack([fx, fy, 0.3 * np.sin(fx / 37.0) * np.cos(fy / 29.0)], axis=1)
    side = rng.integers(0, 4, n_wall)
    t, h = rng.random(n_wall) * L, rng.random(n_wall) * H
    zero, full = np.zeros(n_wall), np.full(n_wall, L)
    wx = np.select([side == 0, side == 1, side == 2, side == 3], [zero, full, t, t])
    wy = np.select([side == 0, side == 1, side == 2, side == 3], [t, t, zero, full])
    walls = np.stack([wx, wy, h], axis=1)
    pts = np.concatenate([floor, walls], axis=0)
    return _records_mm(pts[rng.permutation(n)])


def write_pcd(path, cloud, binary=True):
    """PCD v0.7, FIELDS x y z rgb (rgb packed as float, as PCL writes PointXYZRGB)."""
    n = cloud.shape[0]
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\n"
           "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n"
           % (n, n, "binary" if binary else "ascii"))
    rec = np.empty((n, 4), np.float32)
    rec[:, 0:3] = cloud[:, 0:3]
    rec[:, 3] = cloud[:, 4] if cloud.shape[1] > 4 else 0
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            f.write(rec.tobytes())
        else:
            for r in rec:
                f.write(("%.9g %.9g %.9g %.9g\n" % (r[0], r[1], r[2], r[3])).encode())


def read_pcd(path):
    """Reads ascii/binary PCD with float x y z [rgb|rgba] fields -> (N, 8) float32 PointXYZRGB."""
    with open(path, "rb") as f:
        raw = f.read()
    fields, sizes, types, counts, npts, data_kind = [], [], [], [], None, None
    pos = 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, rest = line.partition(" ")
        if key == "FIELDS":
            fields = rest.split()
        elif key == "SIZE":
            sizes = [int(v) for v in rest.split()]
        elif key == "TYPE":
            types = rest.split()
        elif key == "COUNT":
            counts = [int(v) for v in rest.split()]
        elif key == "POINTS":
            npts = int(rest)
        elif key == "DATA":
            data_kind = rest.strip()
            break
    if not counts:
        counts = [1] * len(fields)
    out = np.zeros((npts, POINT_STRIDE_FLOATS), np.float32)
    out[:, 3] = 1.0
    if data_kind == "ascii":
        arr = np.loadtxt(raw[pos:].decode().splitlines(), dtype=np.float64, ndmin=2)
        col = 0
        for name, cnt in zip(fields, counts):
            if name in ("x", "y", "z"):
                out[:, "xyz".index(name)] = arr[:, col].astype(np.float32)
            elif name in ("rgb", "rgba"):
                out[:, 4] = arr[:, col].astype(np.float32)
            col += cnt
    elif data_kind == "binary":
        dt = []
        for name, sz, ty, cnt in zip(fields, sizes, types, counts):
            base = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 4): "<u4", ("I", 4): "<i4", ("U", 1): "u1",
                    ("U", 2): "<u2", ("I", 2): "<i2", ("I", 1): "i1"}[(ty, sz)]
            dt.append((name, base, (cnt,)) if cnt > 1 else (name, base))
        rec = np.frombuffer(raw, dtype=np.dtype(dt), count=npts, offset=pos)
        for i, name in enumerate("xyz"):
            out[:, i] = rec[name].astype(np.float32)
        for name in ("rgb", "rgba"):
            if name in rec.dtype.names:
                v = rec[name]
                out[:, 4] = v if v.dtype == np.float32 else v.astype(np.uint32).view(np.float32)
    else:
        raise ValueError("unsupported PCD DATA kind: %r" % data_kind)
    return out


def even_planes(cloud, S):
    """S evenly spaced plane x positions strictly inside the x-extent (float32), cfg2/cfg3 style."""
    x = cloud[:, 0]
    lo, hi = np.float32(np.nanmin(x)), np.float32(np.nanmax(x))
    step = (np.float64(hi) - np.float64(lo)) / S
    return (np.float64(lo) + step * (np.arange(S) + 0.5)).astype(np.float32)
