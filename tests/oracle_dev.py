"""An oracle-backed stand-in for api.Cloud -- TEST INFRASTRUCTURE.  Lets the host logic of
polishpathplanning_b200/reference_api.py (plane sweeps, Area2Cloud / compute_boundary / bisection batching,
way-point sampling) run on a machine without a GPU: reference_api.path_generater(..., backend=OracleDev)."""
import numpy as np

from oracle import ppp_oracle as po


class OracleDev:
    def __init__(self, cloud):
        self.oc = po.OracleCloud(cloud)
        self.n = cloud.shape[0]
        self.calls = {}          # method -> number of calls (the batching tests count device round trips)

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def close(self):
        self.oc.close()

    def dev_index(self, *a):
        pass

    def bbox(self):
        return self.oc.minmax()

    def knn(self, k, queries=None, want_d2=True):
        self._count("knn")
        return self.oc.knn(k, queries=queries, want_d2=want_d2)

    def normals_radius(self, r, **kw):
        self._count("normals_radius")
        o, _ = self.oc.normals(radius=r)
        out = np.zeros((self.n, 8), np.float32)
        out[:, 0:3], out[:, 4] = o[:, 0:3], o[:, 3]
        return out

    def principal_curvatures(self, normals, queries, k):
        self._count("principal_curvatures")
        return self.oc.principal_curvatures(normals, queries, k)

    def coverage_mark(self, queries, radius, flags=None):
        self._count("coverage_mark")
        return self.oc.coverage_mark(queries, radius, flags)

    def coverage_mark_radii(self, queries, radii, flags=None):
        self._count("coverage_mark_radii")
        if flags is None:
            flags = np.zeros(self.n, np.uint8)
        for q, r in zip(np.asarray(queries, np.float32), np.asarray(radii, np.float64)):
            if np.isfinite(r) and np.float32(r * r) > 0:
                self.oc.coverage_mark(q[None, :], abs(float(r)), flags)
        return flags

    def slice_bands(self, planes, half_width=2.0, truncate_center=True):
        return self.oc.slice_bands(planes, half_width, truncate_center)

    def slice_contours(self, planes, mode, half_width=2.0, truncate_center=True, **kw):
        self._count("slice_contours")
        if not isinstance(mode, str):
            mode = "A" if mode == 0 else "B"
        return self.oc.slice_contours(planes, mode, half_width, truncate_center)

    def insert_point(self, indices, plane_x, mode):
        if not isinstance(mode, str):
            mode = "A" if mode == 0 else "B"
        y, x, z, _, _ = self.oc.insert_point(indices, float(plane_x), mode)
        return y, x, z

    def sor_mean_distances(self, mean_k=50, sqrt_float=False):
        return self.oc.sor_mean_distances(mean_k, sqrt_float)
