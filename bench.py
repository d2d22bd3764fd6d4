#!/usr/bin/env python
"""bench.py — points/sec through kNN(k=16) + PCA normals + plane slicing + ordered contours.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 1M-point synthetic freeform panel, k=16 normals + 200
slices (5 mm apart, band +-2 mm, SectPath pairing), per GPU.  For N > 1 the panel grows to N x 1M
points (weak scaling), is cut into N x-slabs with a 12 mm halo (polishpathplanning_b200/parallel.py);
every rank runs the whole path on its slab and its planes, results are gathered to rank 0 (NCCL).

One "step" = one pass of the hot path over the cloud, starting from the raw pcl::PointXYZRGB
records: pack + bounding box, grid index build, fused kNN + normals (writes neighbour ids and
normals), band extraction for all planes, per-slice pairing + interpolation + ordering.
  value : device-resident (raw records already in HBM, results left in HBM), CUDA-event timed.
  e2e   : the same step through the host-pointer C ABI (pinned host buffers; H2D of the records,
          D2H of normals and contour nodes inside the timed region).
  roofline : dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json.
  cpu_baseline / --impl reference : the CPU oracle (port of the reference's PCL/FLANN path; the
          reference itself cannot be built here) on the box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "points/sec kNN(k=16)+normals+slicing"
UNIT = "points/s"
N_PER_GPU = 1_000_000
K_NEIGH = 16
S_PER_MILLION = 200       # 5 mm plane spacing on the 1000 mm panel
S_FIXED_TOTAL = None      # --config cfg3: fixed plane count instead
CONFIG_NAME, SCALING = "cfg2", "weak"
HALF_WIDTH = 2.0
HALO_MM = 12.0
PAIRING = "B"             # SectPath::insert_point (src/contour_alg.cpp:165-237)
CPU_SAMPLE_N = 1_000_000  # CPU arm: the full cfg2 workload (seconds of CPU work per step)
CPU_SAMPLE_S = 200


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.proc = None
        self.thr = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.samples.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for p in self.samples:
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_planes(x_min, x_max, S):
    step = (float(x_max) - float(x_min)) / S
    return (float(x_min) + step * (np.arange(S) + 0.5)).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's CPU path)
# ------------------------------------------------------------------------------------------------
def cpu_step(po, cloud, planes, threads):
    oc = po.OracleCloud(cloud)                      # kd-tree build  (Set_kdtree)
    oc.normals(k=K_NEIGH, threads=threads)          # estimate_normal (k-search variant)
    oc.slice_contours(planes, PAIRING, HALF_WIDTH, True, threads=threads)  # rangedX_index + insert_point + ordering
    oc.close()


def cpu_baseline(threads_all=True, reps=1):
    from oracle import ppp_oracle as po
    from polishpathplanning_b200 import synth
    cloud = synth.panel(CPU_SAMPLE_N, seed=0)
    planes = make_planes(cloud[:, 0].min(), cloud[:, 0].max(), CPU_SAMPLE_S)
    cores = po.num_threads() if threads_all else 1
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_step(po, cloud, planes, cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return CPU_SAMPLE_N / best, cores, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ppp_oracle as po
    from polishpathplanning_b200 import synth
    cloud = synth.panel(CPU_SAMPLE_N, seed=0)
    planes = make_planes(cloud[:, 0].min(), cloud[:, 0].max(), CPU_SAMPLE_S)
    cores = po.num_threads()
    for _ in range(args.warmup):
        cpu_step(po, cloud, planes, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(po, cloud, planes, cores)
    dt = time.perf_counter() - t0
    v = CPU_SAMPLE_N * args.steps / dt
    sample = ("oracle port of the reference CPU path on the full workload: %d-point panel (seed 0), "
              "k=%d normals + %d slices, pairing %s, OpenMP over points/slices"
              % (CPU_SAMPLE_N, K_NEIGH, CPU_SAMPLE_S, PAIRING))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def total_slices(gpus):
    return S_FIXED_TOTAL if S_FIXED_TOTAL else int(round(S_PER_MILLION * np.sqrt(gpus)))


def workload_config(gpus):
    if CONFIG_NAME == "cfg3":
        text = ("cfg3: %d-point synthetic freeform panel in total (%d per GPU, seed 0, 1 pt/mm^2), k=%d normals + %d evenly "
                "spaced slices (+-2 mm bands, SectPath pairing)" % (N_PER_GPU * gpus, N_PER_GPU, K_NEIGH, S_FIXED_TOTAL))
    else:
        text = ("cfg2: %d-point synthetic freeform panel per GPU (seed 0, 1 pt/mm^2), k=%d normals + %d slices "
                "per million points (5 mm spacing, +-2 mm bands, SectPath pairing)" % (N_PER_GPU, K_NEIGH, S_PER_MILLION))
    return {"workload": text,
            "points_per_gpu": N_PER_GPU, "k": K_NEIGH, "slices_total": total_slices(gpus),
            "pairing": "B(SectPath)", "partition": "x-slabs+%gmm halo" % HALO_MM if gpus > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB write + 256 MiB read)"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from polishpathplanning_b200 import api, parallel, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks (python -m torch.distributed.run ...)" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=dev)

    ctx = api.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- data (host prep is untimed: it stands for loading the PCD) ----
    n_total = N_PER_GPU * world
    S_total = total_slices(world)
    cloud_g = synth.panel(n_total, seed=0)
    planes_g = make_planes(cloud_g[:, 0].min(), cloud_g[:, 0].max(), S_total)
    if world > 1:
        cuts = parallel.slab_cuts(cloud_g[:, 0], world)
        local_idx, owned = parallel.slab_select(cloud_g, cuts, rank, HALO_MM)
        cloud = np.ascontiguousarray(cloud_g[local_idx])
        my_planes_pos = parallel.owned_planes(planes_g, cuts, rank)
        planes = np.ascontiguousarray(planes_g[my_planes_pos])
        n_owned = int(owned.sum())
    else:
        cloud, planes, owned, local_idx = cloud_g, planes_g, None, None
        n_owned = n_total
    del cloud_g
    n_local = cloud.shape[0]

    with torch.cuda.stream(stream):
        raw_d = torch.from_numpy(cloud).to(dev)
        nstride = 8 if world == 1 else 4       # pcl::Normal records; compact {nx,ny,nz,curv} when gathered over NCCL
        normals_d = torch.empty((n_local, nstride), dtype=torch.float32, device=dev)
        idx_d = torch.empty((n_local, K_NEIGH), dtype=torch.int32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush_rd = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
        owned_d = torch.from_numpy(np.nonzero(owned)[0]).to(dev) if owned is not None else None
    stream.synchronize()

    last = {}
    gat = {}
    # how the normals reach rank 0: "peer" = direct NVLink stores from the search kernel (default),
    # "nccl" = all-gather after the kernel, "none" = left on the owning GPU
    gather_mode = os.environ.get("PPP_BENCH_GATHER", "none" if os.environ.get("PPP_BENCH_NOGATHER") else "peer")
    if world > 1:
        # static gather buffers: equal-sized NCCL gathers, no per-step size exchange
        # The whole local normal array (owned + halo rows) is gathered: rank 0 keeps the owned rows
        # through the static local->global index map, so no per-step row selection kernel is needed
        # (torch.index_select on 1M x 4 floats measured 0.5 ms, ten times the NCCL transfer).
        t = torch.tensor([n_local], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gat["own_cap"] = int(t.item())
        if gather_mode == "peer":
            # rank 0 owns the global result arrays; every rank's kernels store into them over NVLink
            # as they finish (normals through the local->global row map, halo rows skipped; contour
            # nodes + per-slice offsets into the rank's region) and a stream-ordered flag per rank
            # closes the step: no collective on the data path (parallel.PeerSink).
            t = torch.tensor([len(planes)], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            try:
                gat["sink"] = parallel.PeerSink(ctx, dist, dev, rank, world, n_total, node_cap=max(65536, gat["own_cap"] // 4),
                                                S_cap=int(t.item()))
                row_map = np.where(owned, local_idx, -1).astype(np.int32)
                with torch.cuda.stream(stream):
                    gat["row_map"] = torch.from_numpy(row_map).to(dev)
                gat["step"] = 0
            except RuntimeError as e:     # raised on every rank together: no CUDA IPC between these GPUs
                if rank == 0:
                    sys.stderr.write("bench.py: %s; using the NCCL all-gather instead\n" % e)
                gather_mode = "nccl"
        if gather_mode != "peer":
            with torch.cuda.stream(stream):
                gat["own"] = torch.zeros((gat["own_cap"], 4), dtype=torch.float32, device=dev)
                normals_d = gat["own"][:n_local]
                # all-gather (NCCL ring / NVLS collective over NVSwitch) instead of a rooted gather: the
                # rooted gather is a set of point-to-point send/recv pairs and measured ~8x slower here
                gat["own_recv"] = torch.empty((world * gat["own_cap"], 4), dtype=torch.float32, device=dev)

    def ensure_node_buffers(tn):
        if "nodes" in gat and gat["node_cap"] >= tn:
            return
        t = torch.tensor([int(tn * 1.25) + 1024], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gat["node_cap"] = int(t.item())
        with torch.cuda.stream(stream):
            gat["nodes"] = torch.zeros((3, gat["node_cap"] + 1), dtype=torch.float64, device=dev)  # [:,0] = node count
            gat["nodes_recv"] = torch.empty((world * 3, gat["node_cap"] + 1), dtype=torch.float64, device=dev)

    do_gather = gather_mode != "none"

    def dev_step(gather=True):
        gather = gather and do_gather
        c = api.Cloud(ctx, device_ptr=raw_d.data_ptr(), n=n_local, stride_bytes=32)
        if world > 1 and gather_mode == "peer":
            sink = gat["sink"]
            c.dev_set_normal_row_map(gat["row_map"].data_ptr())
            c.dev_normals_knn(K_NEIGH, sink.normals_ptr, 16, idx_ptr=idx_d.data_ptr())
            sink.attach(c)
            res = c.dev_slice_contours(planes, PAIRING, HALF_WIDTH, True)
            if res["total_nodes"] > sink.node_cap:
                raise SystemExit("bench.py: %d contour nodes exceed the peer region (%d)" % (res["total_nodes"], sink.node_cap))
            gat["step"] += 1
            sink.delivered(gat["step"])
            last["nodes"] = res["total_nodes"]
            last["members"] = res["total_members"]
            c.close()
            return
        else:
            c.dev_normals_knn(K_NEIGH, normals_d.data_ptr(), nstride * 4, idx_ptr=idx_d.data_ptr())
        if world > 1 and gather and gather_mode == "nccl":
            # results to rank 0 (the reference's Spline / path connection run on the host of rank 0):
            # owned normals in original index order; issued before the slicing so the NVLink
            # transfer (NCCL's own stream) overlaps the band / contour kernels
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gat["own_recv"], gat["own"])
        if world > 1 and gather and "nodes" in gat:
            nb = gat["nodes"]   # rows y, x, z; column 0 = node count: the library writes the nodes in place
            esz, cap = nb.element_size(), gat["node_cap"]
            c.dev_set_contour_buffers(nb[0].data_ptr() + esz, nb[1].data_ptr() + esz, nb[2].data_ptr() + esz, cap)
        res = c.dev_slice_contours(planes, PAIRING, HALF_WIDTH, True)
        last["nodes"] = res["total_nodes"]
        last["members"] = res["total_members"]
        if world > 1 and gather:
            # contour nodes (y, x, z) of the owned planes
            tn = res["total_nodes"]
            had = "nodes" in gat and gat["node_cap"] >= tn
            ensure_node_buffers(tn)
            with torch.cuda.stream(stream):
                nb = gat["nodes"]
                nb[:, 0] = float(tn)
                if tn and not had:   # first step (or growth): copy from the cloud-owned buffers
                    for j, key in enumerate(("y", "x", "z")):
                        nb[j, 1:tn + 1] = _wrap_f64(torch, res[key], tn, dev)
                dist.all_gather_into_tensor(gat["nodes_recv"], nb)
        c.close()

    def sync_all():
        stream.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def flush_l2():
        # write 256 MiB (evicts everything), then read another 256 MiB so that the lines left in
        # L2 are clean: otherwise the first timed kernel pays for writing the flush buffer back
        with torch.cuda.stream(stream):
            flush.zero_()
            flush_rd.sum()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        flush_l2(); dev_step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.timer_read(0, reset=True)
    l0 = ctx.launch_count()
    sync_all()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2()
        stream.synchronize()
        ctx.timer_begin(0)
        dev_step()
        ctx.timer_end(0)
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    total_ms, regions = ctx.timer_read(0, reset=True)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    assert regions == args.steps
    own_ms = total_ms
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        cnt = torch.tensor([n_owned], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        units = int(cnt.item())
    else:
        units = n_owned
    value = units * args.steps / (total_ms * 1e-3)

    # ---- per-kernel pass (roofline of the dominant kernel), same steps, CUDA events per launch ----
    ctx.kernel_profile(True)
    ctx.kernel_profile_read(reset=True)
    for _ in range(args.steps):
        flush_l2()
        dev_step()
    sync_all()
    prof = ctx.kernel_profile_read(reset=True)
    ctx.kernel_profile(False)
    if os.environ.get("PPP_BENCH_VERBOSE"):
        sys.stderr.write("[rank %d] own step %.4f ms; n_local %d; kernels %s\n" % (
            rank, own_ms / args.steps, n_local,
            {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:4]}))

    # ---- end to end through the host-pointer C ABI (pinned buffers) ----
    pin_cloud = ctx.pinned_empty(cloud.shape, np.float32)
    pin_cloud[...] = cloud
    pin_normals = ctx.pinned_empty((n_local, 8), np.float32)
    node_cap = int(max(last.get("nodes", 0), 1024) * 1.25)
    pin_nodes = tuple(ctx.pinned_empty((node_cap,), np.float64) for _ in range(3))
    e2e_bytes = {}

    def host_step():
        # the reference's call sequence: constructor/Set_kdtree (upload + index), getMinMax3D + plane
        # positions on the host, then estimate_normal + the sweep (one combined C-ABI call)
        c = api.Cloud(ctx, pin_cloud)
        mn, mx = c.bbox()
        pl = planes if world > 1 else make_planes(mn[0], mx[0], S_total)
        _, off, y, x, z = c.normals_and_contours(pl, PAIRING, k=K_NEIGH, half_width=HALF_WIDTH, truncate_center=True,
                                                 normals_out=pin_normals, nodes_out=pin_nodes)
        e2e_bytes["h2d"] = pin_cloud.nbytes + planes.nbytes * 5 + 4 * len(planes)
        e2e_bytes["d2h"] = pin_normals.nbytes + off.nbytes + 3 * y.nbytes
        c.close()
        return off, y, x, z

    for _ in range(min(args.warmup, 3)):
        host_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = units * args.steps / e2e_s

    # ---- untimed check of the peer-store delivery: rank 0's global arrays vs every rank's own results ----
    sink_mismatch = None
    if world > 1 and gather_mode == "peer":
        sync_all()
        sink = gat["sink"]
        S_list = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(S_list, torch.tensor([len(planes)], dtype=torch.int64, device=dev))
        full = torch.empty((n_total, 4), dtype=torch.float32, device=dev)
        node_sum = torch.zeros((world, 4), dtype=torch.float64, device=dev)    # per rank: count, sum y, sum x, sum z
        if rank == 0:
            normals_g, per_rank = sink.read([int(v.item()) for v in S_list])
            full.copy_(torch.from_numpy(normals_g))
            for r, (off, yy, xx, zz) in enumerate(per_rank):
                node_sum[r] = torch.tensor([float(off[-1]), yy.sum(), xx.sum(), zz.sum()], dtype=torch.float64)
        dist.broadcast(full, 0)
        dist.broadcast(node_sum, 0)
        c = api.Cloud(ctx, device_ptr=raw_d.data_ptr(), n=n_local, stride_bytes=32)
        with torch.cuda.stream(stream):
            mine = torch.empty((n_local, 4), dtype=torch.float32, device=dev)
        stream.synchronize()
        c.dev_normals_knn(K_NEIGH, mine.data_ptr(), 16, idx_ptr=idx_d.data_ptr())
        res = c.dev_slice_contours(planes, PAIRING, HALF_WIDTH, True)
        stream.synchronize()
        tn = res["total_nodes"]
        loc = [float(tn)] + [float(ctx.download(res[k], (tn,), np.float64).sum()) for k in ("y", "x", "z")]
        c.close()
        got = full[torch.from_numpy(local_idx[owned]).to(dev)].view(torch.int32)
        bad = (got != mine[owned_d].view(torch.int32)).sum().to(torch.int64)
        bad += int(not np.array_equal(np.asarray(loc), node_sum[rank].cpu().numpy()))
        dist.all_reduce(bad)
        sink_mismatch = int(bad.item())
        del full, mine, got
        sink.close()

    if rank == 0:
        peak, peak_src = read_peaks()
        dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else ("none", (0.0, 0))
        dom_name, (dom_ms, dom_launches) = dom
        alg_bytes = {
            "knn_normals": n_local * (32 + 4 * K_NEIGH),   # 16 B point in + 4k B ids + 16 B normal out
            "knn": n_local * (16 + 4 * K_NEIGH),
            "contour": 16 * last.get("members", 0) + 24 * last.get("nodes", 0),
            "band_count": 16 * n_local, "band_fill": 16 * n_local + 4 * last.get("members", 0),
            "pack_bbox": 12 * n_local + 16 * n_local,
            "cell_count": 16 * n_local, "cell_scatter": 16 * n_local + 20 * n_local,
        }.get(dom_name)
        roof = None
        if alg_bytes and dom_launches:
            achieved = alg_bytes / (dom_ms * 1e-3 / dom_launches) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
            if os.path.exists(tp):
                try:
                    with open(tp) as f:
                        traffic = json.load(f).get(dom_name)
                except Exception:
                    traffic = None
            roof = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": dom_ms / dom_launches,
                    "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
        cpu = None
        if world == 1 and CONFIG_NAME == "cfg2":
            v_all, cores, secs = cpu_baseline(True)
            v_one, _, secs1 = cpu_baseline(False)
            cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "oracle port of the reference CPU path on the full workload (%d-point panel, "
                             "k=%d normals + %d slices pairing %s); %.2f s with %d threads, %.2f s single-thread"
                             % (CPU_SAMPLE_N, K_NEIGH, CPU_SAMPLE_S, PAIRING, secs, cores, secs1),
                   "single_thread_value": v_one}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_bytes.get("h2d", 0)),
                    "d2h_bytes_per_step": int(e2e_bytes.get("d2h", 0)), "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "points_local": n_local, "band_members": last.get("members"), "contour_nodes": last.get("nodes"),
        }
        if world > 1:
            line["config"]["normals_to_rank0"] = {"peer": "NVLink stores from the search kernel into rank 0's array",
                                                  "nccl": "all_gather after the kernel", "none": "not gathered"}[gather_mode]
            if sink_mismatch is not None:
                line["peer_store_mismatches"] = sink_mismatch
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def _wrap_f64(torch, ptr, n, dev):
    """View n doubles at raw device pointer `ptr` as a torch tensor (no copy)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3"],
                    help="cfg2 (default, the driver's contract): 1M points per GPU, k=16, 200*sqrt(N) planes, weak scaling. "
                         "cfg3 (extra, SURVEY 8d): 10M points in TOTAL split over the GPUs, k=32, 1000 planes.")
    args = ap.parse_args()
    if args.config == "cfg3":
        global N_PER_GPU, K_NEIGH, S_PER_MILLION, CONFIG_NAME, SCALING, S_FIXED_TOTAL
        N_PER_GPU = 10_000_000 // max(args.gpus, 1)
        K_NEIGH = 32
        S_FIXED_TOTAL = 1000
        CONFIG_NAME, SCALING = "cfg3", "strong"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
