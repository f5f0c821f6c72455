import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import workloads, libtike.cufft as pt
import contextlib, io
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = workloads.synth_angles(1, 4 * nd, 4 * nd, nd, nd, 32, 1)
psi, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
with pt.CGPtychoSolver(1024, nd, nd, 1, 4 * nd, 4 * nd) as s1:
    data = (s1.fwd(psi, scan, probe[:, 0].contiguous()).abs() ** 2).contiguous()
    psi1 = torch.ones_like(psi)
    g = torch.zeros_like(psi)
    s1._grad(0, psi1, scan, probe, 0, data, None, 1.0, 1.0, 1.0, 0, g)
    print("PTX_STRIP", os.environ.get("PTX_STRIP", "1"), "grad norm %.9e sum %.9e" % (float(torch.linalg.norm(g)), float(g.abs().sum())))
    with contextlib.redirect_stdout(io.StringIO()):
        res = s1.run(data, psi1, scan.clone(), probe.clone(), piter=6, recover_prb=True)
    print("history", s1.history)
    print("ls passes", len(s1.ls_log), "psi norm %.9e" % float(torch.linalg.norm(res["psi"])))
