"""Node-by-node transcription of the reference's per-path host loops -- TEST CHECKER.

compute_transform / Area2Cloud / compute_boundary / bisection / dynamic_adjust_path
(src/Path_Generation.cpp:362-634) and getPath's way-point loop (src/contour_alg.cpp:483-540), written the way
the reference runs them: one path node at a time, one kd-tree query per node, scalar float32 / float64
arithmetic.  `dev` is any object with api.Cloud's query methods (the oracle stand-in or the GPU cloud); the
tests require the BATCHED implementations in polishpathplanning_b200/reference_api.py to give the same
results with the same `dev`, and count how many device calls each needed."""
import numpy as np

f32 = np.float32
DEPTH, ADJUST_THRESHOLD, TOOLTHICKNESS = 0.005, 1.0, 10.0


def compute_transform(dev, normals, point, k=10):
    out, nn0 = dev.principal_curvatures(normals, np.asarray(point, f32).reshape(1, 3), k)
    cur = out[0, 0:3]
    nrm = normals[nn0[0], 0:3]
    cross = np.asarray([nrm[1] * cur[2] - nrm[2] * cur[1], nrm[2] * cur[0] - nrm[0] * cur[2], nrm[0] * cur[1] - nrm[1] * cur[0]], f32)
    Y = np.zeros((4, 4), f32)
    Y[:3, 0], Y[:3, 1], Y[:3, 2], Y[:3, 3], Y[3, 3] = cross, cur, nrm, np.asarray(point, f32), 1.0
    return Y, out[0, 3:5]


def area2cloud(dev, normals, flags, point, key, tool_radius):
    sp = np.asarray(point, np.float64).astype(f32)
    T, pc = compute_transform(dev, normals, sp)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv1, inv0 = float(f32(1) / pc[1]), float(f32(1) / pc[0])
        if pc[0] >= 0 and pc[1] >= 0:
            la = np.sqrt(inv1 * inv1 - (abs(inv1) - DEPTH) ** 2)
            la = tool_radius if la > tool_radius else la
            sa = np.sqrt(inv0 * inv0 - (abs(inv0) - DEPTH) ** 2)
            sa = tool_radius if sa > tool_radius else sa
        else:
            la = abs(inv1) - np.sqrt(inv1 * inv1 - tool_radius * tool_radius)
            la = TOOLTHICKNESS if la > TOOLTHICKNESS else la
            sa = abs(inv0) - np.sqrt(inv0 * inv0 - tool_radius * tool_radius)
            sa = TOOLTHICKNESS if sa > TOOLTHICKNESS else sa
    ell = []
    angle = f32(0.0)
    while angle <= f32(360.0):
        rad = angle * f32(0.017453293)
        x, y, z = f32(la * float(np.cos(rad))), f32(sa * float(np.sin(rad))), f32(0)
        with np.errstate(invalid="ignore"):
            ell.append([x * T[r, 0] + (y * T[r, 1] + (z * T[r, 2] + T[r, 3])) for r in range(3)])
        angle = f32(angle + f32(0.5))
    ell = np.asarray(ell, f32)
    if np.isnan(ell[:, 0]).any():
        return np.full(3, np.nan, f32)
    hi, lo = 0, 0
    for i in range(1, len(ell)):          # std::max_element / min_element: the first extreme element
        if ell[hi, 0] < ell[i, 0]:
            hi = i
        if ell[i, 0] < ell[lo, 0]:
            lo = i
    if not key:
        lan = float((ell[lo, 0] - ell[hi, 0]) / f32(2))
        dev.coverage_mark(np.asarray(point, np.float64).astype(f32).reshape(1, 3), abs(lan), flags) if f32(lan * lan) > 0 else None
        return ell[hi].copy()
    return ell[lo].copy()


def compute_boundary(dev, normals, flags, path, tool_radius, Spline):
    miny, maxy = path.miny(), path.bigy()
    node = {}
    dy = miny + 2
    point, bound, any_node = None, None, False
    while dy < maxy - 2:
        point = path.point(dy)[0]
        dy += tool_radius / 4
        bound = area2cloud(dev, normals, flags, point, False, tool_radius)
        any_node = True
        if np.isnan(bound[0]):
            continue
        node[float(bound[1])] = (float(bound[0]), float(bound[2]))
    if not any_node:
        return None
    bound = area2cloud(dev, normals, flags, point, False, tool_radius)
    node[float(bound[1])] = (float(bound[0]), float(bound[2]))
    keys = sorted(kk for kk in node if not np.isnan(kk)) + ([float("nan")] if np.isnan(bound[1]) else [])
    if len(keys) <= 2:
        return None
    n = len(keys)
    px, py, pz = np.empty(n + 2), np.empty(n + 2), np.empty(n + 2)
    for i, kk in enumerate(keys):
        px[i + 1], py[i + 1], pz[i + 1] = node[kk][0], kk, node[kk][1]
    px[0], py[0], pz[0] = px[1], py[1] - 20, pz[1]
    px[n + 1], py[n + 1], pz[n + 1] = px[n], py[n] + 20, pz[n]
    return Spline(py, px, pz)


def bisection(dev, normals, flags, node, boundary, tool_radius, itr=0):
    node = np.array(node, np.float64)
    if itr > 5:
        return node
    ab = area2cloud(dev, normals, flags, node, True, tool_radius)
    with np.errstate(invalid="ignore"):
        if ab[1] < boundary.miny() or ab[1] > boundary.bigy():
            return node
    if np.isnan(ab[1]):
        node[0] = np.nan
        return node
    bp = boundary.point(float(ab[1]))[0]
    n0 = float(ab[0]) - bp[0]
    if abs(n0) < ADJUST_THRESHOLD:
        return node
    node[0] = node[0] - n0
    if np.isnan(node[0]):
        return node
    return bisection(dev, normals, flags, node, boundary, tool_radius, itr + 1)


def dynamic_adjust_path(dev, cloud, normals, flags, origin_path, pre_path, tool_radius, Spline):
    boundary = compute_boundary(dev, normals, flags, pre_path, tool_radius, Spline)
    if boundary is None:
        return None
    miny, maxy = origin_path.miny(), origin_path.bigy()
    num = int((maxy - miny) / 5)
    new_path = {}
    for i in range(1, num):
        dy = ((maxy - miny) / num * i) + miny
        node = bisection(dev, normals, flags, origin_path.point(dy)[0], boundary, tool_radius)
        idx, _ = dev.knn(3, queries=np.asarray(node, np.float64).astype(f32).reshape(1, 3), want_d2=False)
        p = cloud[idx[0, 0]]
        new_path[float(p[1])] = (float(p[0]), float(p[2]))
    keys = sorted(new_path)
    return Spline(np.asarray(keys), np.asarray([new_path[kk][0] for kk in keys]), np.asarray([new_path[kk][1] for kk in keys]))


def getpath_waypoints(dev, normals, splines, resolution, inv):
    """The two loops of SectPath::getPath, way-point by way-point."""
    lists, flag = [], 1
    for path in splines[1:-1]:
        one = []
        dy = path.miny() + 5
        while dy < path.bigy() - 5:
            p = path.point(dy)[0]
            w = np.asarray([p[0], p[1], p[2], 1.0], f32)
            one.append((((inv[:, 0] * w[0] + inv[:, 1] * w[1]) + inv[:, 2] * w[2]) + inv[:, 3] * w[3]).astype(f32))
            dy += resolution
        if flag == -1:
            one.reverse()
        lists.append(one)
        flag *= -1
    xyz, ids, rots, tails = [], [], [], []
    for one in lists:
        for w in one:
            idx, _ = dev.knn(1, queries=w[None, :3].copy(), want_d2=False)
            N = normals[idx[0, 0], 0:3]
            A = (-N).astype(f32)
            O = np.asarray([A[1] * f32(0) - A[2] * f32(0), A[2] * f32(1) - A[0] * f32(0), A[0] * f32(0) - A[1] * f32(1)], f32)
            Nn = np.asarray([O[1] * A[2] - O[2] * A[1], O[2] * A[0] - O[0] * A[2], O[0] * A[1] - O[1] * A[0]], f32)
            xyz.append(w[:3]); ids.append(idx[0, 0]); rots.append(np.stack([Nn, O, A], axis=1))
        tails.append(len(xyz) - 1)
    return np.asarray(xyz, f32).reshape(-1, 3), np.asarray(ids, np.int32), np.asarray(rots, f32).reshape(-1, 3, 3), np.asarray(tails, np.int64)
