"""Python mirror of the reference's operator interface for the hot path (same names, argument
meaning and error behaviour), backed by libppp_gpu.so.  The C++ mirror that a reference
maintainer links is polishpathplanning_b200/host/ (Path_Generate_gpu.h, contour_alg_gpu.h).

  path_generater  : include/Path_Generate.h:35-76  + src/Path_Generation.cpp   (gen-2, ./main)
  SectPath        : include/contour_alg.h:50-91    + src/contour_alg.cpp       (config.txt flow)

Only the hot-path members are mirrored (estimate_normal, Set_kdtree, rangedX_index, insert_point,
slicing_method, path_track / OnePath, the GenPath plane sweep); visualisation, dynamic adjustment,
way-point export and the GSL spline stay with the reference's own host code.
Plane positions are generated here exactly as each reference loop does, in float32.
"""
import numpy as np

from . import api, synth

_f32 = np.float32


# -------------------------------------------------------------------------------------------------
# plane sweeps (host logic).  All arithmetic in float32 like the reference.
# -------------------------------------------------------------------------------------------------
def planes_gen2_contact(min_x, max_x, tool_radius):
    """Contact_Path_Generation: float locateX = min.x + toolRadius (double add, then float);
    locateX += int(2R) while locateX < max.x.  src/Path_Generation.cpp:711-723"""
    step = int(tool_radius * 2)
    out = []
    loc = _f32(float(_f32(min_x)) + float(tool_radius))
    while loc < _f32(max_x) and step > 0:
        out.append(loc)
        loc = _f32(loc + _f32(step))
    return np.asarray(out, _f32)


def planes_gen2_slicing(min_x, max_x, tool_radius):
    """slicing_method: min.x += step/2 (INTEGER division), += step.  src/Path_Generation.cpp:295-303"""
    step = int(tool_radius * 2)
    out = []
    m = _f32(_f32(min_x) + _f32(step // 2))
    while m < _f32(max_x) and step > 0:
        out.append(m)
        m = _f32(m + _f32(step))
    return np.asarray(out, _f32)


def planes_gen1_slicing(min_x, max_x):
    """free slicing_method: first plane min.x + 40, step 40.  src/slicing_method.cpp:617-626"""
    out = []
    m = _f32(_f32(min_x) + _f32(40))
    while m < _f32(max_x):
        out.append(m)
        m = _f32(m + _f32(40))
    return np.asarray(out, _f32)


def planes_sectpath(min_x, max_x, tool_radius):
    """SectPath::GenPath: backward from centre - step while > min.x (inserted at the front), then
    forward from the centre while < max.x.  Returned in Path_set order.  src/contour_alg.cpp:305-327"""
    step = int(tool_radius * 2)
    centre = _f32(_f32(_f32(min_x) + _f32(max_x)) / _f32(2))
    front, back = [], []
    loc = _f32(centre - _f32(step))
    while loc > _f32(min_x) and step > 0:
        front.insert(0, loc)
        loc = _f32(loc - _f32(step))
    loc = centre
    while loc < _f32(max_x) and step > 0:
        back.append(loc)
        loc = _f32(loc + _f32(step))
    return np.asarray(front + back, _f32)


class Spline:
    """include/Spline.h:7-51 — two GSL Steffen splines x(y), z(y) over the ordered contour nodes.
    gsl_interp_steffen [upstream, recalled: GSL is not in the image]: monotone cubic Hermite of
    Steffen (1990); same arithmetic order as the oracle restatement, in float64."""

    def __init__(self, point_y, point_x, point_z):
        y = np.ascontiguousarray(point_y, np.float64)
        if y.shape[0] < 3 or not np.all(np.diff(y) > 0):
            raise ValueError("Spline needs >= 3 strictly increasing y (GSL would abort)")
        self.y = y
        self._cx = self._coeffs(y, np.ascontiguousarray(point_x, np.float64))
        self._cz = self._coeffs(y, np.ascontiguousarray(point_z, np.float64))
        self.small_y, self.big_y = float(y[0]), float(y[-1])

    @staticmethod
    def _coeffs(xa, ya):
        n = xa.shape[0]
        h = xa[1:] - xa[:-1]
        s = (ya[1:] - ya[:-1]) / h
        yp = np.empty(n, np.float64)
        yp[0] = s[0]
        hi, him1, si, sim1 = h[1:], h[:-1], s[1:], s[:-1]
        pi = (sim1 * hi + si * him1) / (him1 + hi)
        yp[1:-1] = (np.copysign(1.0, sim1) + np.copysign(1.0, si)) * np.minimum(np.abs(sim1), np.minimum(np.abs(si), 0.5 * np.abs(pi)))
        yp[-1] = s[-1]
        a = (yp[:-1] + yp[1:] - 2 * s) / h / h
        b = (3 * s - 2 * yp[:-1] - yp[1:]) / h
        return a, b, yp[:-1].copy(), ya[:-1].copy()

    def _eval(self, c, yq):
        a, b, cc, d = c
        yq = np.asarray(yq, np.float64)
        i = np.clip(np.searchsorted(self.y, yq, side="right") - 1, 0, self.y.shape[0] - 2)
        dx = yq - self.y[i]
        out = d[i] + dx * (cc[i] + dx * (b[i] + dx * a[i]))
        return np.where((yq >= self.y[0]) & (yq <= self.y[-1]), out, np.nan)

    def point(self, y):
        """Eigen::Vector3d point(double y): (x(y), y, z(y)); vectorised over y."""
        y = np.atleast_1d(np.asarray(y, np.float64))
        return np.stack([self._eval(self._cx, y), y, self._eval(self._cz, y)], axis=1)

    def miny(self):
        return self.small_y

    def bigy(self):
        return self.big_y


class _Base:
    NORMAL_RADIUS = 2.5   # normal_estimation.setRadiusSearch(2.5): src/Path_Generation.cpp:329
    BAND_HALF_WIDTH = 2.0  # setFilterLimits(-2 + position, 2 + position): src/Path_Generation.cpp:100

    def __init__(self, ctx=None, device=0):
        self._ctx = ctx or api.Context(device)
        self._gpu = None
        self.cloud = None
        self.Path_set = []

    def _load(self, cloud_name, change_range=True):
        if isinstance(cloud_name, str):
            try:
                cloud = synth.read_pcd(cloud_name)
            except (OSError, ValueError) as e:
                # the reference prints PCL_ERROR and carries on with an empty cloud
                print("Cloudn't read file! (%s)" % e)
                cloud = np.zeros((0, synth.POINT_STRIDE_FLOATS), np.float32)
        else:
            cloud = np.ascontiguousarray(cloud_name, np.float32).copy()
        if change_range:
            synth.reference_ctor_scale(cloud)
        self.cloud = cloud
        self._invalidate()

    def _invalidate(self):
        """xyz changed (voxel_down / trans2center / smooth / remove_outlier in the reference):
        drop the device copy and its index."""
        if self._gpu is not None:
            self._gpu.close()
        self._gpu = None

    def _dev(self):
        if self._gpu is None:
            self._gpu = api.Cloud(self._ctx, self.cloud)
        return self._gpu

    # -- mirrored members ---------------------------------------------------------------------------
    def Set_kdtree(self):
        """kdtree.setInputCloud(cloud): builds the device index."""
        self._dev().dev_index(16, 0.0)

    def estimate_normal(self):
        """NormalEstimation, radius 2.5, viewpoint (0,0,0) -> (N, 8) pcl::Normal records."""
        self.cloud_with_normals = self._dev().normals_radius(self.NORMAL_RADIUS)
        return self.cloud_with_normals

    def rangedX_index(self, position):
        """std::vector<int> rangedX_index(int position): the argument is truncated to int."""
        position = int(position)
        off, idx = self._dev().slice_bands(np.asarray([position], _f32), self.BAND_HALF_WIDTH, True)
        return idx

    def getMinMax3D(self):
        return self._dev().bbox()

    def _contours(self, planes, mode):
        return self._dev().slice_contours(np.asarray(planes, _f32), mode, self.BAND_HALF_WIDTH, True)

    # -- "next" rows: consumers of the ordered contours --------------------------------
    def splines(self):
        """Spline objects of Path_set (paths with < 3 nodes are skipped, GSL would abort on them)."""
        return [Spline(y, x, z) for (y, x, z) in self.Path_set if len(y) >= 3]

    def drawpath_samples(self, path, gen2=True):
        """The sample points drawpath snaps to the cloud: gen-2 takes 200 samples
        dy = (maxy-miny)/200*i + miny (src/Path_Generation.cpp:644-646); SectPath steps dy += 1
        while dy < maxy (src/contour_alg.cpp:273-283)."""
        miny, maxy = path.miny(), path.bigy()
        if gen2:
            dy = (maxy - miny) / 200 * np.arange(200) + miny
        else:
            dy = miny + np.arange(int(np.ceil(maxy - miny)) + 1, dtype=np.float64)
            dy = dy[dy < maxy]
        return path.point(dy)

    def drawpath(self, path, gen2=True):
        """Indices of the cloud points drawpath recolours: nearestKSearch(point, k)[0] per sample
        (k = 3 in gen-2, 1 in SectPath: only element [0] is used)."""
        pts = self.drawpath_samples(path, gen2).astype(np.float32)   # PCLp.x = pathPoint[0] (double -> float)
        idx, _ = self._dev().knn(1, queries=np.ascontiguousarray(pts), want_d2=False)
        return idx[:, 0]

    def compute_transform(self, points, k=10):
        """Batched compute_transform (src/Path_Generation.cpp:362-400; k = 50 in the gen-3 planner):
        returns (pt2Base matrices (n,4,4) float32, principle_curvature (n,2))."""
        pts = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1, 3))
        out, nn0 = self._dev().principal_curvatures(self.cloud_with_normals, pts, k)
        nrm = self.cloud_with_normals[nn0, 0:3]
        cur = out[:, 0:3]
        # Eigen cross: normalVector.cross(curvatureVector), float32
        cr = np.stack([nrm[:, 1] * cur[:, 2] - nrm[:, 2] * cur[:, 1], nrm[:, 2] * cur[:, 0] - nrm[:, 0] * cur[:, 2],
                       nrm[:, 0] * cur[:, 1] - nrm[:, 1] * cur[:, 0]], axis=1).astype(np.float32)
        Y = np.zeros((pts.shape[0], 4, 4), np.float32)
        Y[:, :3, 0], Y[:, :3, 1], Y[:, :3, 2], Y[:, :3, 3], Y[:, 3, 3] = cr, cur, nrm, pts, 1.0
        return Y, out[:, 3:5]

    def compute_coverage(self, nodes, radius):
        """compute_coverage for a batch of nodes (src/Path_Generation.cpp:483-496)."""
        if not hasattr(self, "coverage_flag") or self.coverage_flag.shape[0] != self.cloud.shape[0]:
            self.coverage_flag = np.zeros(self.cloud.shape[0], np.uint8)
        q = np.ascontiguousarray(np.asarray(nodes, np.float64).astype(np.float32).reshape(-1, 3))
        self._dev().coverage_mark(q, radius, self.coverage_flag)
        return self.coverage_flag

    def get_coverage(self):
        """get_coverage (src/Path_Generation.cpp:757-771): rate = yes / (yes + no), float arithmetic."""
        yes = _f32(int(self.coverage_flag.sum()))
        no = _f32(int(self.coverage_flag.shape[0] - self.coverage_flag.sum()))
        return float(yes / (yes + no))


class path_generater(_Base):
    """gen-2 planner (what ./main builds): brute-force pairing with greedy flags (variant A)."""

    def __init__(self, cloud_name, Radius, ctx=None, device=0):
        super().__init__(ctx, device)
        self.toolRadius = float(Radius)
        self._load(cloud_name, change_range=True)

    def insert_point(self, indices, PlanePoint):
        """std::map<double, std::vector<double>> insert_point(indices, PlanePoint): returns the map
        as (y, x, z) arrays in ascending y.  `indices`: strictly ascending point indices (every reference
        call site passes rangedX_index(PlanePoint[0]), src/Path_Generation.cpp:300-301,665)."""
        return self._dev().insert_point(np.asarray(indices, np.int32), _f32(PlanePoint[0]), api.PPP_PAIR_GEN2)

    def path_track(self, plane_point):
        y, x, z = self.insert_point(self.rangedX_index(plane_point[0]), plane_point)
        self.Path_set.append((y, x, z))  # Spline(node_number, point_y, point_x, point_z)

    def slicing_method(self):
        """All planes of the sweep in one device pass; returns (planes, node_offsets, y, x, z)."""
        mn, mx = self.getMinMax3D()
        planes = planes_gen2_slicing(mn[0], mx[0], self.toolRadius)
        off, y, x, z = self._contours(planes, api.PPP_PAIR_GEN2)
        print("number of paths: %d" % len(planes))
        return planes, off, y, x, z

    def Contact_Path_Generation(self, adjust=False):
        """Plane sweep + path_track of Contact_Path_Generation (dynamic adjustment is out of scope)."""
        if adjust:
            raise NotImplementedError("dynamic_adjust_path is sequential host logic outside the hot path")
        mn, mx = self.getMinMax3D()
        planes = planes_gen2_contact(mn[0], mx[0], self.toolRadius)
        off, y, x, z = self._contours(planes, api.PPP_PAIR_GEN2)
        self.Path_set = [(y[off[s]:off[s + 1]], x[off[s]:off[s + 1]], z[off[s]:off[s + 1]]) for s in range(len(planes))]
        print("Number of paths: %d" % len(planes))
        return planes


def sor_select(distances, n_valid, stddev_mul=1.0, negative=False):
    """Second pass of pcl::StatisticalOutlierRemoval::applyFilterIndices: running double sums in index
    order over all distances (np.cumsum accumulates sequentially), variance over n_valid - 1, inliers
    distance <= mean + stddev_mul * stddev.  Returns (kept indices ascending, threshold)."""
    d = np.ascontiguousarray(distances, np.float32)
    if d.shape[0] == 0 or n_valid < 1:
        return np.zeros(0, np.int32), float("nan")
    total = float(np.cumsum(d.astype(np.float64))[-1])
    sq_total = float(np.cumsum((d * d).astype(np.float64))[-1])        # float32 products, widened when added
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = np.float64(total) / np.float64(n_valid)
        variance = (np.float64(sq_total) - np.float64(total) * np.float64(total) / np.float64(n_valid)) / (np.float64(n_valid) - 1.0)
        thr = float(mean + np.float64(stddev_mul) * np.sqrt(variance))
    outlier = (d <= thr) if negative else (d > thr)
    return np.flatnonzero(~outlier).astype(np.int32), thr


class SectPath(_Base):
    """SectPath (contour_alg.h): kd-tree pairing without flags (variant B), centre-out sweep."""

    def __init__(self, cloud_name, Tool_Radius, ChangeRange=True, RemoveOutlier=False, ctx=None, device=0):
        super().__init__(ctx, device)
        self.toolRadius = float(Tool_Radius)
        self._load(cloud_name, change_range=ChangeRange)
        if RemoveOutlier and self.cloud.shape[0] > 50:      # src/contour_alg.cpp:29
            self.remove_outlier()

    def insert_point(self, indices, PlanePoint):
        return self._dev().insert_point(np.asarray(indices, np.int32), _f32(PlanePoint[0]), api.PPP_PAIR_SECT)

    def remove_outlier(self, mean_k=50, stddev_mul=1.0):
        """SectPath::remove_outlier (src/contour_alg.cpp:101-108): StatisticalOutlierRemoval with
        setMeanK(50), setStddevMulThresh(1.0), filtered in place.  The kNN pass runs on the device; the
        mean / stddev / threshold pass is the reference's sequential double arithmetic."""
        dist, n_valid = self._dev().sor_mean_distances(mean_k)
        keep, _ = sor_select(dist, n_valid, stddev_mul)
        self.cloud = np.ascontiguousarray(self.cloud[keep])
        self._invalidate()
        return keep

    def OnePath(self, plane_point):
        return self.insert_point(self.rangedX_index(plane_point[0]), plane_point)

    def GenPath(self):
        mn, mx = self.getMinMax3D()
        planes = planes_sectpath(mn[0], mx[0], self.toolRadius)
        off, y, x, z = self._contours(planes, api.PPP_PAIR_SECT)
        self.Path_set = [(y[off[s]:off[s + 1]], x[off[s]:off[s + 1]], z[off[s]:off[s + 1]]) for s in range(len(planes))]
        print("Number of paths: %d" % len(planes))
        print("Number of Point Cloud: %d" % self.cloud.shape[0])
        return planes
