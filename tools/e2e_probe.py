"""e2e leg of bench.py (host arrays in, host gradient out) for several chunk sizes (ptheta)."""
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
    data = slv.fwd(psi, scan, probe[:, 0]).abs().square_().contiguous().cpu().numpy()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h = {"data": pin(data), "psi": pin(np.ones_like(w["psi"])), "scan": pin(w["scan"]), "probe": pin(w["probe"])}
nbytes = sum(v.nbytes for v in h.values())
for pth in (1, 2, 4, 8):
    with pt.CGPtychoSolver(S, N, N, pth, nz, n) as s1:
        for _ in range(2):
            s1.grad_ptycho_batch(h["data"], h["psi"], h["scan"], h["probe"])
        t0 = time.perf_counter()
        for _ in range(10):
            s1.grad_ptycho_batch(h["data"], h["psi"], h["scan"], h["probe"])
        dt = (time.perf_counter() - t0) / 10
    print("ptheta %d: %.2f ms per step, %.0f patterns/s, H2D %.1f GB/s" % (pth, dt * 1e3, T * S / dt, nbytes / dt / 1e9))
import cProfile, pstats
with pt.CGPtychoSolver(S, N, N, 1, nz, n) as s1:
    for _ in range(2):
        s1.grad_ptycho_batch(h["data"], h["psi"], h["scan"], h["probe"])
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        s1.grad_ptycho_batch(h["data"], h["psi"], h["scan"], h["probe"])
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(12)
