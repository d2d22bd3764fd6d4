"""A SECOND restatement of PCL 1.10's point-normal arithmetic, in vectorised numpy float32 -- TEST CHECKER.

Written separately from oracle/ppp_oracle.cpp (different language, array-at-a-time instead of
point-at-a-time, numpy's own float32 kernels for sqrt / arctan2 / cos / sin instead of glibc's) from the
published algorithm [upstream, recalled: PCL 1.10.0 common/impl/centroid.hpp
computeMeanAndCovarianceMatrix, features/impl/normal_3d.hpp solvePlaneParameters +
flipNormalTowardsViewpoint, common/impl/eigen.hpp computeRoots / computeRoots2 / eigen33]; the oracle tests
require the two to agree.  Every operation is a float32 numpy ufunc on float32 arrays, so each step rounds
to float32 exactly once, as the C++ does with -ffp-contract=off."""
import numpy as np

F = np.float32


def normals_from_lists(P, idx, viewpoint=(0.0, 0.0, 0.0)):
    """P: (N, 3) float32 points; idx: (N, k) int32 neighbour lists in (d2, index) order (no padding).
    Returns (N, 4) float32: nx, ny, nz, curvature."""
    P = np.asarray(P, F)
    n, k = idx.shape
    acc = np.zeros((9, n), F)
    for j in range(k):                                  # sequential accumulation in list order
        q = P[idx[:, j]]
        x, y, z = q[:, 0], q[:, 1], q[:, 2]
        for a, v in enumerate((x * x, x * y, x * z, y * y, y * z, z * z, x, y, z)):
            acc[a] = acc[a] + v
    acc = acc / F(k)
    c00 = acc[0] - acc[6] * acc[6]
    c01 = acc[1] - acc[6] * acc[7]
    c02 = acc[2] - acc[6] * acc[8]
    c11 = acc[3] - acc[7] * acc[7]
    c12 = acc[4] - acc[7] * acc[8]
    c22 = acc[5] - acc[8] * acc[8]
    cov = [c00, c01, c02, c11, c12, c22]
    scale = np.zeros(n, F)
    for c in cov:
        scale = np.maximum(scale, np.abs(c))
    scale = np.where(scale <= np.finfo(F).tiny, F(1.0), scale)
    m00, m01, m02, m11, m12, m22 = (c / scale for c in cov)
    # computeRoots
    c0 = m00 * m11 * m22 + F(2.0) * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01
    c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12
    c2 = m00 + m11 + m22

    def roots2(b, c):
        d = (b * b).astype(np.float64) - 4.0 * c.astype(np.float64)     # Scalar d = Scalar (b * b - 4.0 * c)
        d = np.maximum(d.astype(F), F(0.0))
        sd = np.sqrt(d)
        return np.zeros_like(b), F(0.5) * (b - sd), F(0.5) * (b + sd)

    inv3 = F(1.0 / 3.0)
    sqrt3 = np.sqrt(F(3.0))
    c2_3 = c2 * inv3
    a_3 = np.minimum((c1 - c2 * c2_3) * inv3, F(0.0))
    half_b = F(0.5) * (c0 + c2_3 * (F(2.0) * c2_3 * c2_3 - c1))
    q = np.minimum(half_b * half_b + a_3 * a_3 * a_3, F(0.0))
    rho = np.sqrt(-a_3)
    theta = np.arctan2(np.sqrt(-q), half_b) * inv3
    ct, st = np.cos(theta), np.sin(theta)
    r0 = c2_3 + F(2.0) * rho * ct
    r1 = c2_3 - rho * (ct + sqrt3 * st)
    r2 = c2_3 - rho * (ct - sqrt3 * st)
    swap = r0 >= r1
    r0, r1 = np.where(swap, r1, r0), np.where(swap, r0, r1)
    swap = r1 >= r2
    r1n, r2n = np.where(swap, r2, r1), np.where(swap, r1, r2)
    swap2 = swap & (r0 >= r1n)
    r0, r1 = np.where(swap2, r1n, r0), np.where(swap2, r0, r1n)
    r2 = r2n
    q0, q1, q2 = roots2(c2, c1)
    use2 = (np.abs(c0) < np.finfo(F).eps) | (r0 <= F(0.0))
    root0 = np.where(use2, q0, r0)
    eigenvalue = root0 * scale
    d0, d1, d2 = m00 - root0, m11 - root0, m22 - root0

    def cross(a, b):
        return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])

    rows = ((d0, m01, m02), (m01, d1, m12), (m02, m12, d2))
    v1, v2, v3 = cross(rows[0], rows[1]), cross(rows[0], rows[2]), cross(rows[1], rows[2])
    sq = lambda v: v[0] * v[0] + (v[1] * v[1] + v[2] * v[2])      # Eigen's 3-vector squaredNorm association
    l1, l2, l3 = sq(v1), sq(v2), sq(v3)
    pick1 = (l1 >= l2) & (l1 >= l3)
    pick2 = ~pick1 & (l2 >= l1) & (l2 >= l3)
    v = [np.where(pick1, a, np.where(pick2, b, c)) for a, b, c in zip(v1, v2, v3)]
    ln = np.sqrt(np.where(pick1, l1, np.where(pick2, l2, l3)))
    nx, ny, nz = v[0] / ln, v[1] / ln, v[2] / ln
    tr = (c00 + c11) + c22
    with np.errstate(divide="ignore", invalid="ignore"):
        curv = np.where(tr != 0, np.abs(eigenvalue / tr), F(0.0)).astype(F)
    vp = np.asarray(viewpoint, F)
    w = [vp[i] - P[:, i] for i in range(3)]
    flip = ((w[0] * nx + w[1] * ny) + w[2] * nz) < 0
    nx, ny, nz = np.where(flip, -nx, nx), np.where(flip, -ny, ny), np.where(flip, -nz, nz)
    return np.stack([nx, ny, nz, curv], axis=1).astype(F)
