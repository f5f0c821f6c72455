"""Host -> device and device -> host copy rate from pinned memory, alone or on every GPU of the box at once:
the ceiling of the host-array legs of bench.py (`e2e`).
    python tools/pcie_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py"""
import os

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("gloo")
x = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (64, 512):
    xs, ds = x[: n * 1024 * 1024 // 4], d[: n * 1024 * 1024 // 4]
    ds.copy_(xs, non_blocking=True)
    torch.cuda.synchronize()
    for name, dst, src in (("H2D", ds, xs), ("D2H", xs, ds)):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e1.record()
        e1.synchronize()
        rate = torch.tensor([20 * xs.numel() * 4 / e0.elapsed_time(e1) / 1e6], dtype=torch.float64)
        if world > 1:
            rates = [torch.zeros_like(rate) for _ in range(world)]
            dist.all_gather(rates, rate)
        else:
            rates = [rate]
        if rank == 0:
            r = [float(v) for v in rates]
            print("%s %d MiB pinned, %d GPU(s) at once: per GPU %s GB/s, sum %.1f GB/s"
                  % (name, n, world, " ".join("%.1f" % v for v in r), sum(r)))
