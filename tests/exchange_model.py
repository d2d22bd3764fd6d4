"""numpy restatement of the device exchange (polishpathplanning_b200/csrc/exchange.cu) -- TEST CHECKER.

Same arithmetic, step by step: float32 min/max of the finite x, a 1024-bin histogram over the global
range in double, equal-count cuts in integer arithmetic, owner by bin, halo copies by comparing x with
the cut positions, destination order = ascending global index.  The GPU tests compare the slabs the
kernels deliver with this; the gloo tests use it as the exchange stand-in on CPU."""
import numpy as np

BINS = 1024


def exchange_model(chunks, starts, halo):
    """chunks[r]: float32 (n_r, >=3) records of rank r (file records [starts[r], starts[r+1])).
    Returns (slabs, info): slabs[d] = (xyz float32 (n_local, 3), w int32 (n_local,)) with w = global
    index for owned rows and ~index for halo copies; info = dict(cuts, x_range, n_owned[d])."""
    world = len(chunks)
    pts = np.concatenate([np.asarray(c, np.float32)[:, :3] for c in chunks], axis=0)
    n = pts.shape[0]
    assert n == int(starts[-1])
    fin = np.isfinite(pts).all(axis=1)
    if fin.any():
        gmin, gmax = float(pts[fin, 0].min()), float(pts[fin, 0].max())
    else:
        gmin = gmax = 0.0
    span = max(gmax - gmin, 1e-9)
    inv = float(BINS) / span
    b = np.floor((pts[:, 0].astype(np.float64) - gmin) * inv)
    b = np.clip(np.where(fin, b, 0), 0, BINS - 1).astype(np.int64)
    hist = np.bincount(b[fin], minlength=BINS).astype(np.int64)
    cum = np.cumsum(hist)
    total = int(cum[-1])
    cutbin = np.zeros(world + 1, np.int64)
    cutbin[world] = BINS
    for r in range(1, world):
        k = int(np.argmax(cum * world >= total * r)) if np.any(cum * world >= total * r) else BINS - 1
        cutbin[r] = k + 1
    cutx = np.empty(world + 1, np.float64)
    cutx[0], cutx[world] = -np.inf, np.inf
    for r in range(1, world):
        cutx[r] = gmin + span * float(cutbin[r]) / float(BINS)
    owner = np.zeros(n, np.int64)
    for r in range(1, world):
        owner += (b >= cutbin[r])
    owner[~fin] = 0
    xd = pts[:, 0].astype(np.float64)
    slabs, n_owned = [], []
    gidx = np.arange(n, dtype=np.int64)
    for d in range(world):
        own = owner == d
        with np.errstate(invalid="ignore"):
            lower = fin & (owner > d) & (xd < cutx[d + 1] + halo)   # owner above d: every slab in between also matches
            upper = fin & (owner < d) & (xd >= cutx[d] - halo)
        sel = own | lower | upper
        idx = gidx[sel]
        w = np.where(own[sel], idx, ~idx).astype(np.int32)
        slabs.append((pts[sel].copy(), w))
        n_owned.append(int(own.sum()))
    return slabs, {"cuts": cutx, "x_range": np.asarray([gmin, gmax]), "n_owned": n_owned}
