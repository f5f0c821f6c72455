import torch, time
x = torch.empty(512*1024*1024//4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (64, 512):
    xs = x[: n*1024*1024//4]; ds = d[: n*1024*1024//4]
    ds.copy_(xs, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ds.copy_(xs, non_blocking=True)
    e1.record(); e1.synchronize()
    print("H2D %d MiB pinned: %.1f GB/s" % (n, 10 * xs.numel() * 4 / e0.elapsed_time(e1) / 1e6))
    e0.record()
    for _ in range(10): xs.copy_(ds, non_blocking=True)
    e1.record(); e1.synchronize()
    print("D2H %d MiB pinned: %.1f GB/s" % (n, 10 * xs.numel() * 4 / e0.elapsed_time(e1) / 1e6))
