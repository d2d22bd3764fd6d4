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
    bad = owned & np.isfinite(reach) & np.isfinite(x)   # non-finite points have no neighbours at all
    bad &= ((x - reach) < lo) & np.isfinite(lo) | ((x + reach) >= hi) & np.isfinite(hi)
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


# -------------------------------------------------------------------------------------------------
# Device-side redistribution: the exchange step of the multi-GPU path (NCCL all-to-all-v).
# -------------------------------------------------------------------------------------------------
IDX_COL = 5    # pcl::PointXYZRGB padding floats carry the global point index (as int32 bits) ...
OWNED_COL = 6  # ... and the "owned by this rank" flag while a record travels between ranks


def redistribute(dist, chunk, global_start, rank, world, halo, bins=4096):
    """Spatial redistribution of a cloud that starts out split by ORIGINAL INDEX (rank r holds the
    records [global_start, global_start + len(chunk)) of the file, in order), e.g. each rank read
    its share of the PCD.  Afterwards rank r holds every point of its x-slab plus the halo copies
    from its neighbours, still in ascending global index order (so index tie-breaks are those of
    the single-GPU run).

    chunk: torch float32 (n_r, 8) PointXYZRGB records on the rank's device (or CPU with gloo).
    Collectives: all_reduce (x range, 4096-bin x histogram -> equal-count cuts), all_to_all_single
    (counts, then the records: the halo exchange).  Returns (local_records, global_idx int64,
    owned bool, cuts float64 tensor on CPU)."""
    import torch
    dev = chunk.device
    n = chunk.shape[0]
    x = chunk[:, 0]
    fin = torch.isfinite(chunk[:, 0]) & torch.isfinite(chunk[:, 1]) & torch.isfinite(chunk[:, 2])
    big = torch.finfo(torch.float32).max
    lo = torch.where(fin, x, torch.full_like(x, big)).min() if n else torch.tensor(big, device=dev)
    hi = torch.where(fin, x, torch.full_like(x, -big)).max() if n else torch.tensor(-big, device=dev)
    rng = torch.stack([lo, -hi]).to(torch.float64)
    dist.all_reduce(rng, op=dist.ReduceOp.MIN)
    x_min, x_max = float(rng[0]), float(-rng[1])
    span = max(x_max - x_min, 1e-9)
    # equal-count cuts from a global histogram (identical on every rank)
    b = torch.clamp(((x.to(torch.float64) - x_min) / span * bins).floor().to(torch.int64), 0, bins - 1)
    hist = torch.bincount(b[fin], minlength=bins).to(torch.int64)
    dist.all_reduce(hist)
    cum = torch.cumsum(hist, 0).cpu().numpy()
    total = int(cum[-1])
    cuts = [-np.inf]
    for r in range(1, world):
        k = int(np.searchsorted(cum, total * r / world, side="left"))
        cuts.append(x_min + span * (k + 1) / bins)
    cuts.append(np.inf)
    cuts = np.asarray(cuts, np.float64)
    cuts_t = torch.tensor(cuts[1:-1], dtype=torch.float64, device=dev)
    x64 = x.to(torch.float64)
    owner = torch.bucketize(x64, cuts_t, right=True)           # cut[r] <= x < cut[r+1]
    owner = torch.where(fin, owner, torch.zeros_like(owner))      # non-finite points stay with rank 0
    rec = chunk.clone()
    gidx = torch.arange(global_start, global_start + n, device=dev, dtype=torch.int32)
    rec[:, IDX_COL] = gidx.view(torch.float32)
    send_parts, counts = [], []
    cl = torch.tensor(cuts, dtype=torch.float64, device=dev)
    for d in range(world):
        own = owner == d
        near = fin & ~own & (x64 >= cl[d] - halo) & (x64 < cl[d + 1] + halo)
        sel = own | near
        part = rec[sel]
        part[:, OWNED_COL] = own[sel].to(torch.float32)
        send_parts.append(part)
        counts.append(part.shape[0])
    send = torch.cat(send_parts, 0) if send_parts else rec[:0]
    cnt_s = torch.tensor(counts, dtype=torch.int64, device=dev)
    cnt_r = torch.empty_like(cnt_s)
    dist.all_to_all_single(cnt_r, cnt_s)
    rc = [int(v) for v in cnt_r.cpu()]
    recv = torch.empty((sum(rc), chunk.shape[1]), dtype=chunk.dtype, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=counts)
    g = recv[:, IDX_COL].contiguous().view(torch.int32).to(torch.int64)
    owned = recv[:, OWNED_COL] > 0.5
    local = recv.clone()
    local[:, IDX_COL] = 0.0
    local[:, OWNED_COL] = 0.0
    return local, g, owned, cuts


# -------------------------------------------------------------------------------------------------
# Result delivery without a collective: rank 0 owns the global result arrays, every rank's kernels
# store into them over NVLink (CUDA IPC mapping), and a stream-ordered flag per rank tells rank 0
# when a step's results have landed.
# -------------------------------------------------------------------------------------------------
def peer_sink_layout(world, n_total, node_cap, S_cap):
    """Byte layout of the PeerSink buffer (every section 256-byte aligned):
    [normals n_total x 16][rank 0: offsets | y | x | z] ... [rank world-1: ...][flags world x 128]."""
    a256 = lambda v: (int(v) + 255) // 256 * 256
    normals_bytes = a256(int(n_total) * 16)
    off_bytes = a256((int(S_cap) + 1) * 8)
    arr_bytes = a256(int(node_cap) * 8)
    region_bytes = off_bytes + 3 * arr_bytes
    flags_at = normals_bytes + int(world) * region_bytes
    return {"normals_bytes": normals_bytes, "off_bytes": off_bytes, "arr_bytes": arr_bytes, "region_bytes": region_bytes,
            "flags_at": flags_at, "total_bytes": flags_at + int(world) * 128}


class PeerSink:
    """Global result arrays on rank 0, written in place by all ranks.

    Layout of the one peer buffer (rank 0's HBM):
      normals   n_total x 16 B   {nx, ny, nz, curvature} in ORIGINAL index order
      per rank  offsets (S_cap + 1) int64, then y / x / z: node_cap doubles each
      flags     world x 128 B    flag r = number of the last step rank r has delivered

    Every rank passes `normals_ptr` with its local->global row map to the search
    (Cloud.dev_set_normal_row_map + dev_normals_*), calls attach(cloud) before dev_slice_contours and
    delivered(step) after it.  On rank 0, delivered() also makes the stream wait for every peer's flag,
    so work queued behind it (and a host sync) sees all ranks' results of that step.  There is no ack
    back to the writers: a consumer that reads while the next step runs must double-buffer.
    """

    def __init__(self, ctx, dist, device, rank, world, n_total, node_cap, S_cap):
        import torch
        self.ctx, self.rank, self.world = ctx, rank, world
        # one layout for all ranks: capacities are the maxima over the ranks' requests
        caps = torch.tensor([int(n_total), int(node_cap), int(S_cap)], dtype=torch.int64, device=device)
        dist.all_reduce(caps, op=dist.ReduceOp.MAX)
        self.n_total, self.node_cap, self.S_cap = (int(v) for v in caps.tolist())
        lay = peer_sink_layout(world, self.n_total, self.node_cap, self.S_cap)
        self._normals_bytes, self._off_bytes, self._arr_bytes = lay["normals_bytes"], lay["off_bytes"], lay["arr_bytes"]
        self._region_bytes, self._flags_at, total = lay["region_bytes"], lay["flags_at"], lay["total_bytes"]
        hbuf = torch.zeros(64, dtype=torch.uint8, device=device)
        self.base, self._dist = None, dist
        ok, why = 1, ""
        if rank == 0:
            try:
                self.base, handle = ctx.peer_buffer_alloc(total)      # zero-filled
                hbuf.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
            except Exception as e:                                    # noqa: BLE001 - reported to every rank below
                ok, why = 0, str(e)
        dist.broadcast(hbuf, 0)
        if rank != 0:
            try:
                self.base = ctx.peer_buffer_open(hbuf.cpu().numpy().tobytes())
            except Exception as e:                                    # noqa: BLE001
                ok, why = 0, str(e)
        # every rank learns whether ALL mappings exist, so that callers can fall back together
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if self.base is not None:
                (ctx.peer_buffer_free if rank == 0 else ctx.peer_buffer_close)(self.base)
                self.base = None
            raise RuntimeError("peer buffers are not available on every rank" + (": " + why if why else ""))
        self.normals_ptr = self.base

    def region(self, r):
        at = self.base + self._normals_bytes + r * self._region_bytes
        return {"off": at, "y": at + self._off_bytes, "x": at + self._off_bytes + self._arr_bytes,
                "z": at + self._off_bytes + 2 * self._arr_bytes}

    def flag(self, r):
        return self.base + self._flags_at + 128 * r

    def attach(self, cloud):
        """Route the contour nodes + per-slice offsets of `cloud` into this rank's region."""
        g = self.region(self.rank)
        cloud.dev_set_contour_buffers(g["y"], g["x"], g["z"], self.node_cap)
        cloud.dev_set_contour_offsets_buffer(g["off"], self.S_cap + 1)

    def delivered(self, step):
        """Stream-ordered: this rank's results of `step` (1, 2, ...) are in place."""
        self.ctx.signal(self.flag(self.rank), step)
        if self.rank == 0:
            for r in range(1, self.world):
                self.ctx.wait(self.flag(r), step)

    def read(self, S_per_rank):
        """Rank 0, after a host sync: (normals (n_total, 4), [(offsets, y, x, z) per rank])."""
        assert self.rank == 0
        normals = self.ctx.download(self.normals_ptr, (self.n_total, 4), np.float32)
        out = []
        for r in range(self.world):
            g = self.region(r)
            off = self.ctx.download(g["off"], (S_per_rank[r] + 1,), np.int64)
            n = int(off[-1])
            out.append((off, self.ctx.download(g["y"], (n,), np.float64), self.ctx.download(g["x"], (n,), np.float64),
                        self.ctx.download(g["z"], (n,), np.float64)))
        return normals, out

    def close(self):
        """Collective: peers unmap, then rank 0 frees."""
        if self.rank != 0:
            self.ctx.peer_buffer_close(self.base)
        self._dist.barrier()
        if self.rank == 0:
            self.ctx.peer_buffer_free(self.base)
        self.base = None
