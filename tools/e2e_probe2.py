"""Where does the e2e step lose time?  Event-timed H2D / kernel / D2H per chunk in the same 2-stream pipeline."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import libtike.cufft as pt
import workloads
T = 8
w = workloads.c2_single_angle(ntheta=T)
S, N, nz, n = w["nscan"], w["ndet"], w["nz"], w["n"]
with pt.CGPtychoSolver(S, N, N, T, nz, n) as slv:
    psi, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
    data = slv.fwd(psi, scan, probe[:, 0]).abs().square_().contiguous().cpu()
hd = data.pin_memory()
dd = [torch.empty_like(data[:1], device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
torch.cuda.synchronize()
# (1) copies only, alternating streams
for variant in ("copies only, 2 streams", "copies only, 1 stream"):
    t0 = time.perf_counter()
    for rep in range(5):
        for c in range(T):
            st = streams[c % 2] if "2" in variant else streams[0]
            with torch.cuda.stream(st):
                dd[c % 2].copy_(hd[c:c + 1], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("%s: %.2f ms per %d chunks -> %.1f GB/s" % (variant, dt * 1e3, T, hd.numel() * 4 / dt / 1e9))
# (2) copies + a kernel per chunk
with pt.CGPtychoSolver(S, N, N, 1, nz, n) as s1:
    g = [torch.zeros((1, nz, n), dtype=torch.complex64, device="cuda") for _ in range(2)]
    p1, sc1, pr1 = torch.ones_like(psi[:1]), scan[:1].contiguous(), probe[:1].contiguous()
    for variant in ("copy + kernel, 2 streams",):
        t0 = time.perf_counter()
        for rep in range(5):
            for c in range(T):
                st = streams[c % 2]
                with torch.cuda.stream(st):
                    dd[c % 2].copy_(hd[c:c + 1], non_blocking=True)
                    s1._grad(0, p1, sc1, pr1, 0, dd[c % 2], None, 1.0, 1.0, 1.0, 0, g[c % 2])
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print("%s: %.2f ms per %d chunks -> %.1f GB/s" % (variant, dt * 1e3, T, hd.numel() * 4 / dt / 1e9))
