import time, torch, os, sys
import torch.nn.functional as F
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
x = torch.randn(4, 48, 56, 56); w = torch.randn(48, 48, 3, 3)
for nt in (4, 8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    F.conv2d(x, w, padding=1)
    t = time.perf_counter()
    for _ in range(20): F.conv2d(x, w, padding=1)
    print(nt, "threads: conv 48ch 56^2 B4 ms =", (time.perf_counter() - t) / 20 * 1e3, flush=True)
