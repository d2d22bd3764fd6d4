import os, time, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 1_000_000
src = torch.randn(n + 17000, 4, device=dev); idx = torch.arange(n, device=dev)
own = torch.zeros(n, 4, device=dev); recv = torch.empty(world * n, 4, device=dev)
recv_list = [torch.empty_like(own) for _ in range(world)] if rank == 0 else None
s = torch.cuda.Stream(device=dev)
def timeit(name, fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps * 1e3
    if rank == 0: print("%-40s %.3f ms" % (name, dt), flush=True)
with torch.cuda.stream(s):
    timeit("index_select", lambda: torch.index_select(src, 0, idx, out=own))
    timeit("all_gather_into_tensor 16MB", lambda: dist.all_gather_into_tensor(recv, own))
    timeit("gather 16MB", lambda: dist.gather(own, recv_list, dst=0))
    small = torch.zeros(3, 150000, dtype=torch.float64, device=dev); rs = torch.empty(world * 3, 150000, dtype=torch.float64, device=dev)
    timeit("all_gather 3.6MB", lambda: dist.all_gather_into_tensor(rs, small))
    def fill():
        small[:, 0] = 5.0
        small[0, 1:100001] = small[1, 1:100001]
    timeit("slice assigns", fill)
dist.destroy_process_group()
