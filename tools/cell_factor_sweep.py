import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from polishpathplanning_b200 import api, synth
n = 1_000_000
ctx = api.Context(0); dev = torch.device("cuda", 0)
cloud = synth.panel(n, 0); raw = torch.from_numpy(cloud).to(dev)
nrm = torch.empty((n, 8), dtype=torch.float32, device=dev); idx = torch.empty((n, 16), dtype=torch.int32, device=dev)
planes = synth.even_planes(cloud, 200)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    c = api.Cloud(ctx, device_ptr=raw.data_ptr(), n=n, stride_bytes=32)
    c.dev_normals_knn(16, nrm.data_ptr(), 32, idx_ptr=idx.data_ptr())
    c.dev_slice_contours(planes, "B")
    c.close()
for f in ("1.30", "1.35", "1.40", "1.45", "1.50", "1.60"):
    os.environ["PPP_CELL_FACTOR"] = f
    for _ in range(3): step()
    ctx.sync(); ctx.timer_read(2, True)
    for _ in range(20):
        flush.zero_(); torch.cuda.synchronize()
        ctx.timer_begin(2); step(); ctx.timer_end(2)
    ms, k = ctx.timer_read(2, True)
    ctx.kernel_profile(True); ctx.kernel_profile_read(True)
    for _ in range(5): step()
    ctx.sync(); prof = ctx.kernel_profile_read(True); ctx.kernel_profile(False)
    print("f=%s step %.4f ms  knn %.4f redo %.4f" % (f, ms / k, prof["knn_normals"][0] / 5, prof["knn_redo"][0] / 5), flush=True)
