"""Multi-GPU sharding of the hot path: one process per GPU, x-slab decomposition with halo.

The reference is single-process; its path shards naturally because every query
point and every slice is independent given read access to its neighbourhood (§8e).  Each rank
owns the points (and the slicing planes) of one x-interval and additionally holds a halo of
`halo` mm on both sides, so that
  * the k nearest / radius neighbours of every OWNED point, and
  * the +-half_width band of every OWNED plane plus the nearest-neighbour reach of its pairing
are all inside the rank's local cloud: no data-path collective is needed; the only exchange is the
gather of results to rank 0 (normals by original index, contour nodes by plane), which is what
the reference's downstream Spline / path connection consumes on the host.

Exactness guard: `halo_violations` reports owned points whose k-th neighbour distance reaches the
edge of the local x-extent (they would need a wider halo); callers re-run those with a wider halo
(the synthetic panels never trigger it with the default 12 mm).

The functions here are backend-agnostic (numpy in / numpy out) so the world_size-2 gloo tests can
drive them on CPU with the oracle standing in for the CUDA library.
"""
import numpy as np


def slab_cuts(x, world):
    """world+1 cut positions (float64) splitting the finite x values into equal-count intervals.
    cuts[0] = -inf, cuts[-1] = +inf."""
    xs = np.sort(x[np.isfinite(x)].astype(np.float64))
    cuts = [-np.inf]
    for r in range(1, world):
        cuts.append(float(xs[min(len(xs) - 1, (len(xs) * r) // world)]) if len(xs) else 0.0)
    cuts.append(np.inf)
    return np.asarray(cuts, np.float64)


def slab_select(cloud, cuts, rank, halo):
    """Indices (ascending, global numbering) of the rank's local points and the owned mask among them."""
    x = cloud[:, 0].astype(np.float64)
    lo, hi = cuts[rank], cuts[rank + 1]
    local = np.nonzero((x >= lo - halo) & (x < hi + halo))[0]
    xl = x[local]
    owned = (xl >= lo) & (xl < hi)
    return local.astype(np.int64), owned


def owned_planes(planes, cuts, rank):
    """Positions (into `planes`) of the planes this rank owns: cut[rank] <= x < cut[rank+1]."""
    p = np.asarray(planes, np.float64)
    return np.nonzero((p >= cuts[rank]) & (p < cuts[rank + 1]))[0]


def halo_violations(local_cloud, owned, kth_d2, cuts, rank, halo):
    """Owned points whose k-th neighbour distance reaches beyond the halo (local numbering)."""
    x = local_cloud[:, 0].astype(np.float64)
    lo, hi = cuts[rank] - halo, cuts[rank + 1] + halo
    reach = np.sqrt(np.maximum(kth_d2.astype(np.float64), 0.0)) * (1.0 + 1e-6)
    bad = owned & (((x - reach) < lo) & np.isfinite(lo) | ((x + reach) >= hi) & np.isfinite(hi))
    return np.nonzero(bad)[0]


def gather_to_rank0(dist, arrays, rank, world, device=None):
    """Variable-length gather of a list of equally-long 1-D/2-D numpy arrays (or torch tensors on
    `device`) from every rank to rank 0 with torch.distributed.  Returns on rank 0 a list (per
    rank) of lists of numpy arrays; None elsewhere."""
    import torch
    n = int(arrays[0].shape[0])
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device=device))
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    out = [[] for _ in range(world)] if rank == 0 else None
    for a in arrays:
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        if device is not None:
            t = t.to(device)
        pad_shape = (mx,) + tuple(t.shape[1:])
        buf = torch.zeros(pad_shape, dtype=t.dtype, device=device)
        buf[:n] = t
        recv = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, recv, dst=0)
        if rank == 0:
            for r in range(world):
                out[r].append(recv[r][:sizes[r]].cpu().numpy())
    return out


def assemble_normals(n_total, stride_floats, per_rank):
    """per_rank: list of (global_idx, normals_rows). Rows not covered stay NaN."""
    full = np.full((n_total, stride_floats), np.nan, np.float32)
    for gidx, rows in per_rank:
        full[gidx] = rows
    return full


def assemble_contours(S, per_rank):
    """per_rank: list of (plane_positions, node_offsets, y, x, z) for the planes each rank owns.
    Returns global (node_offsets, y, x, z) ordered by plane position."""
    counts = np.zeros(S, np.int64)
    for pos, off, _, _, _ in per_rank:
        counts[pos] = np.diff(off)
    goff = np.zeros(S + 1, np.int64)
    np.cumsum(counts, out=goff[1:])
    y = np.empty(int(goff[-1]), np.float64)
    x = np.empty_like(y)
    z = np.empty_like(y)
    for pos, off, yy, xx, zz in per_rank:
        for j, s in enumerate(pos):
            a, b = int(off[j]), int(off[j + 1])
            y[goff[s]:goff[s + 1]] = yy[a:b]
            x[goff[s]:goff[s + 1]] = xx[a:b]
            z[goff[s]:goff[s + 1]] = zz[a:b]
    return goff, y, x, z
