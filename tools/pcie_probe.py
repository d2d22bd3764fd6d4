#!/usr/bin/env python
"""D2H / H2D copy time vs size on this box (CUDA events, pinned host memory)."""
import torch
dev = torch.device("cuda", 0)
for direction in ("d2h", "h2d"):
    for size in (4 << 10, 64 << 10, 256 << 10, 1 << 20, 4 << 20, 32 << 20):
        d = torch.empty(size, dtype=torch.uint8, device=dev)
        h = torch.empty(size, dtype=torch.uint8).pin_memory()
        for _ in range(3):
            (h if direction == "d2h" else d).copy_(d if direction == "d2h" else h, non_blocking=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            (h if direction == "d2h" else d).copy_(d if direction == "d2h" else h, non_blocking=True)
            b.record(); b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print("%s %9d B: median %.1f us  (%.1f GB/s)" % (direction, size, ts[5], size / ts[5] / 1e3))
