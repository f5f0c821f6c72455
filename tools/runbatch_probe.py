"""Where does run_batch spend its time?  4 angles of the c4 shape: run() on device-resident inputs vs run_batch on
host arrays (piter = 8), wall clock.   usage: python tools/runbatch_probe.py [ndet=256]"""
import contextlib, io, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import workloads, libtike.cufft as pt
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = 4
w = workloads.synth_angles(A, 4 * nd, 4 * nd, nd, nd, 32, 1)
psi_t, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
with pt.CGPtychoSolver(1024, nd, nd, 1, 4 * nd, 4 * nd) as s1, contextlib.redirect_stdout(io.StringIO()):
    data = torch.cat([(s1.fwd(psi_t[t:t + 1], scan[t:t + 1], probe[t:t + 1, 0].contiguous()).abs() ** 2) for t in range(A)])
    psi0 = torch.ones_like(psi_t)
    h = {"data": data.cpu().numpy(), "psi": psi0.cpu().numpy(), "scan": w["scan"], "probe": w["probe"]}
    s1.run(data[:1], psi0[:1], scan[:1].clone(), probe[:1].clone(), piter=2, recover_prb=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(A):
        s1.run(data[t:t + 1], psi0[t:t + 1], scan[t:t + 1].clone(), probe[t:t + 1].clone(), piter=8, recover_prb=True)
    torch.cuda.synchronize(); t_dev = time.perf_counter() - t0
    s1.run_batch(h["data"][:1], h["psi"][:1], h["scan"][:1], h["probe"][:1], piter=2, recover_prb=True)
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s1.run_batch(h["data"], h["psi"], h["scan"], h["probe"], piter=8, recover_prb=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
sys.stderr.write("ndet %d: 4 x run(piter=8) device resident %.1f ms (%.1f angle-it/s); run_batch %s ms (best %.1f angle-it/s)\n" % (
    nd, t_dev * 1e3, 32 / t_dev, ["%.1f" % (x * 1e3) for x in ts], 32 / min(ts)))
