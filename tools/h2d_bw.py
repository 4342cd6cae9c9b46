"""Raw pinned-host <-> device copy bandwidth on this box (context for the e2e number)."""
import time

import torch

n = 1_500_000_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (src, dst) in {"h2d": (h, d), "d2h": (d, h)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print("%s: %.1f GB/s (%.1f ms per 1.5 GB)" % (name, n / dt / 1e9, dt * 1e3))
