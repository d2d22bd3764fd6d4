"""ctypes wrapper around oracle/libppp_oracle.so — CPU ORACLE, TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (polishpathplanning_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libppp_oracle.so")

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(_HERE, "ppp_oracle.cpp")
    if force or not os.path.exists(_SO) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.ppo_cloud_create.restype = C.c_void_p
        L.ppo_cloud_create.argtypes = [_f32p, C.c_int64, C.c_int64]
        L.ppo_cloud_destroy.argtypes = [C.c_void_p]
        L.ppo_minmax.argtypes = [_f32p, C.c_int64, C.c_int64, _f32p, _f32p]
        L.ppo_knn.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, C.c_int, _i32p, _f32p, C.c_int]
        L.ppo_knn_brute.argtypes = [_f32p, C.c_int64, C.c_int64, _f32p, C.c_int64, C.c_int64, C.c_int, _i32p, _f32p]
        L.ppo_radius.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, C.c_double, _i32p, _i64p, _i32p, _f32p, C.c_int]
        L.ppo_normals.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, _f32p, C.c_int, _f32p, _i32p, C.c_int]
        L.ppo_normal_from_list.argtypes = [_f32p, C.c_int64, C.c_int64, _i32p, C.c_int, _f32p, _f32p, C.c_int, _f32p]
        L.ppo_band.restype = C.c_int64
        L.ppo_band.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_int, _i32p]
        L.ppo_slice_bands.argtypes = [_f32p, C.c_int64, C.c_int64, _f32p, C.c_int, C.c_float, C.c_int, _i64p, _i32p, C.c_int]
        L.ppo_insert_point.restype = C.c_int64
        L.ppo_insert_point.argtypes = [C.c_void_p, _i32p, C.c_int64, C.c_float, C.c_int, _f64p, _f64p, _f64p,
                                       C.c_int64, _i32p, _i32p, _i64p, _i64p]
        L.ppo_slice_contours.restype = C.c_int64
        L.ppo_slice_contours.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_int, C.c_int, _i64p, _f64p,
                                         _f64p, _f64p, C.c_int64, C.c_int]
        L.ppo_planes.argtypes = [C.c_int, C.c_float, C.c_float, C.c_double, _f32p, C.c_int]
        L.ppo_num_threads.restype = C.c_int
        L.ppo_steffen_eval.argtypes = [_f64p, _f64p, C.c_int64, _f64p, C.c_int64, _f64p]
        L.ppo_principal_curvatures.argtypes = [C.c_void_p, _f32p, C.c_int64, _f32p, C.c_int64, C.c_int64, C.c_int, _f32p, _i32p]
        L.ppo_sor_mean_distances.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, C.POINTER(C.c_int64), C.c_int]
        L.ppo_sor_select.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_double, C.c_int, _i32p, C.POINTER(C.c_double)]
        L.ppo_sor_select.restype = C.c_int64
        L.ppo_coverage_mark.argtypes = [C.c_void_p, _f32p, C.c_int64, C.c_int64, C.c_double, C.POINTER(C.c_ubyte)]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _ip(a):
    return a.ctypes.data_as(_i32p)


def _lp(a):
    return a.ctypes.data_as(_i64p)


def _dp(a):
    return a.ctypes.data_as(_f64p)


class OracleCloud:
    """Owns a float32 (N, stride) array (stride 8 = pcl::PointXYZRGB, or 3/4) + the oracle kd-tree."""

    def __init__(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        assert pts.ndim == 2 and pts.shape[1] >= 3
        self.pts = pts
        self.n = pts.shape[0]
        self.sf = pts.shape[1]
        self._h = lib().ppo_cloud_create(_fp(pts), self.n, self.sf)

    def close(self):
        if self._h:
            lib().ppo_cloud_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def minmax(self):
        mn = np.zeros(3, np.float32)
        mx = np.zeros(3, np.float32)
        lib().ppo_minmax(_fp(self.pts), self.n, self.sf, _fp(mn), _fp(mx))
        return mn, mx

    def knn(self, k, queries=None, threads=0, want_d2=True):
        if queries is None:
            nq, q, qsf = self.n, None, 0
        else:
            queries = np.ascontiguousarray(queries, np.float32)
            nq, q, qsf = queries.shape[0], _fp(queries), queries.shape[1]
        idx = np.empty((nq, k), np.int32)
        d2 = np.empty((nq, k), np.float32) if want_d2 else None
        lib().ppo_knn(self._h, q, nq, qsf, k, _ip(idx), _fp(d2) if want_d2 else None, threads)
        return idx, d2

    def knn_brute(self, k, queries=None):
        if queries is None:
            nq, q, qsf = self.n, None, 0
        else:
            queries = np.ascontiguousarray(queries, np.float32)
            nq, q, qsf = queries.shape[0], _fp(queries), queries.shape[1]
        idx = np.empty((nq, k), np.int32)
        d2 = np.empty((nq, k), np.float32)
        lib().ppo_knn_brute(_fp(self.pts), self.n, self.sf, q, nq, qsf, k, _ip(idx), _fp(d2))
        return idx, d2

    def radius(self, r, queries=None, threads=0):
        if queries is None:
            nq, q, qsf = self.n, None, 0
        else:
            queries = np.ascontiguousarray(queries, np.float32)
            nq, q, qsf = queries.shape[0], _fp(queries), queries.shape[1]
        counts = np.empty(nq, np.int32)
        lib().ppo_radius(self._h, q, nq, qsf, float(r), _ip(counts), None, None, None, threads)
        offsets = np.zeros(nq + 1, np.int64)
        np.cumsum(counts, out=offsets[1:])
        idx = np.empty(int(offsets[-1]), np.int32)
        d2 = np.empty(int(offsets[-1]), np.float32)
        lib().ppo_radius(self._h, q, nq, qsf, float(r), _ip(counts), _lp(offsets), _ip(idx), _fp(d2), threads)
        return counts, offsets, idx, d2

    def normals(self, radius=None, k=None, viewpoint=(0.0, 0.0, 0.0), cov_variant=0, threads=0):
        assert (radius is None) != (k is None)
        vp = np.asarray(viewpoint, np.float32)
        out = np.empty((self.n, 4), np.float32)
        cnt = np.empty(self.n, np.int32)
        lib().ppo_normals(self._h, 0 if k is None else 1, float(radius or 0.0), int(k or 0), _fp(vp), cov_variant,
                          _fp(out), _ip(cnt), threads)
        return out, cnt

    def band(self, plane_x, half_width=2.0, truncate_center=True):
        n = lib().ppo_band(_fp(self.pts), self.n, self.sf, plane_x, half_width, int(truncate_center), None)
        idx = np.empty(n, np.int32)
        lib().ppo_band(_fp(self.pts), self.n, self.sf, plane_x, half_width, int(truncate_center), _ip(idx))
        return idx

    def slice_bands(self, planes, half_width=2.0, truncate_center=True, threads=0):
        planes = np.ascontiguousarray(planes, np.float32)
        S = planes.shape[0]
        off = np.zeros(S + 1, np.int64)
        lib().ppo_slice_bands(_fp(self.pts), self.n, self.sf, _fp(planes), S, half_width, int(truncate_center),
                              _lp(off), None, threads)
        idx = np.empty(int(off[-1]), np.int32)
        lib().ppo_slice_bands(_fp(self.pts), self.n, self.sf, _fp(planes), S, half_width, int(truncate_center),
                              _lp(off), _ip(idx), threads)
        return off, idx

    def principal_curvatures(self, normals, queries, k):
        normals = np.ascontiguousarray(normals, np.float32)
        queries = np.ascontiguousarray(queries, np.float32)
        out = np.empty((queries.shape[0], 5), np.float32)
        nn0 = np.empty(queries.shape[0], np.int32)
        lib().ppo_principal_curvatures(self._h, _fp(normals), normals.shape[1], _fp(queries), queries.shape[0],
                                       queries.shape[1], int(k), _fp(out), _ip(nn0))
        return out, nn0

    def sor_mean_distances(self, mean_k=50, sqrt_float=False, threads=0):
        """First pass of StatisticalOutlierRemoval. Returns (distances float32[N], n_valid)."""
        dist = np.empty(self.n, np.float32)
        nv = C.c_int64(0)
        rc = lib().ppo_sor_mean_distances(self._h, mean_k, int(sqrt_float), _fp(dist), C.byref(nv), threads)
        if rc != 0:
            raise ValueError("sor: cloud has no more than mean_k points")
        return dist, nv.value

    @staticmethod
    def sor_select(dist, n_valid, std_mul=1.0, negative=False):
        """Second pass: (kept indices, threshold)."""
        dist = np.ascontiguousarray(dist, np.float32)
        kept = np.empty(dist.shape[0], np.int32)
        thr = C.c_double(0)
        n = lib().ppo_sor_select(_fp(dist), dist.shape[0], n_valid, std_mul, int(negative), _ip(kept), C.byref(thr))
        return kept[:n].copy(), thr.value

    def coverage_mark(self, queries, radius, flags=None):
        queries = np.ascontiguousarray(queries, np.float32)
        if flags is None:
            flags = np.zeros(self.n, np.uint8)
        lib().ppo_coverage_mark(self._h, _fp(queries), queries.shape[0], queries.shape[1], float(radius),
                                flags.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return flags

    def insert_point(self, indices, plane_x, mode):
        """mode 'A' (gen-2) or 'B' (SectPath). Returns (y, x, z, left_pair, right_pair)."""
        indices = np.ascontiguousarray(indices, np.int32)
        m = indices.shape[0]
        cap = max(m, 1)
        y = np.empty(cap, np.float64)
        x = np.empty(cap, np.float64)
        z = np.empty(cap, np.float64)
        lp = np.empty(cap, np.int32)
        rp = np.empty(cap, np.int32)
        nl = C.c_int64(0)
        nr = C.c_int64(0)
        n = lib().ppo_insert_point(self._h, _ip(indices), m, plane_x, 0 if mode == "A" else 1, _dp(y), _dp(x), _dp(z),
                                   cap, _ip(lp), _ip(rp), C.byref(nl), C.byref(nr))
        return y[:n].copy(), x[:n].copy(), z[:n].copy(), lp[:nl.value].copy(), rp[:nr.value].copy()

    def slice_contours(self, planes, mode, half_width=2.0, truncate_center=True, threads=0):
        planes = np.ascontiguousarray(planes, np.float32)
        S = planes.shape[0]
        off = np.zeros(S + 1, np.int64)
        cap = max(self.n, 1)
        while True:
            y = np.empty(cap, np.float64)
            x = np.empty(cap, np.float64)
            z = np.empty(cap, np.float64)
            tot = lib().ppo_slice_contours(self._h, _fp(planes), S, half_width, int(truncate_center),
                                           0 if mode == "A" else 1, _lp(off), _dp(y), _dp(x), _dp(z), cap, threads)
            if tot <= cap:
                return off, y[:tot].copy(), x[:tot].copy(), z[:tot].copy()
            cap = int(tot)


def steffen_eval(xa, ya, xq):
    xa = np.ascontiguousarray(xa, np.float64)
    ya = np.ascontiguousarray(ya, np.float64)
    xq = np.ascontiguousarray(xq, np.float64)
    out = np.empty(xq.shape[0], np.float64)
    st = lib().ppo_steffen_eval(_dp(xa), _dp(ya), xa.shape[0], _dp(xq), xq.shape[0], _dp(out))
    if st != 0:
        raise ValueError("steffen: need >= 3 strictly increasing abscissae")
    return out


def normal_from_list(pts, nb, q, viewpoint=(0, 0, 0), cov_variant=0):
    pts = np.ascontiguousarray(pts, np.float32)
    nb = np.ascontiguousarray(nb, np.int32)
    q = np.ascontiguousarray(q, np.float32)
    vp = np.asarray(viewpoint, np.float32)
    out = np.empty(4, np.float32)
    lib().ppo_normal_from_list(_fp(pts), pts.shape[0], pts.shape[1], _ip(nb), nb.shape[0], _fp(q), _fp(vp),
                               cov_variant, _fp(out))
    return out


PLANE_VARIANTS = {"gen2_contact": 0, "gen2_slicing": 1, "gen1_slicing": 2, "sectpath": 3, "gen3_two_thread": 4, "gen3_sdir": 5}


def planes(variant, min_x, max_x, tool_radius):
    v = PLANE_VARIANTS[variant] if isinstance(variant, str) else int(variant)
    out = np.empty(1, np.float32)
    n = lib().ppo_planes(v, float(np.float32(min_x)), float(np.float32(max_x)), float(tool_radius), _fp(out), 0)
    out = np.empty(max(n, 1), np.float32)
    lib().ppo_planes(v, float(np.float32(min_x)), float(np.float32(max_x)), float(tool_radius), _fp(out), n)
    return out[:n]


def num_threads():
    return lib().ppo_num_threads()
