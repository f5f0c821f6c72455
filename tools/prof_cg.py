"""Run a few CG iterations on one c2 / c4 angle: target of a launch-list capture
(ncu --metrics gpu__time_duration.sum) that shows where an iteration's GPU time goes."""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import libtike.cufft as pt  # noqa: E402
import workloads  # noqa: E402

ndet = int(sys.argv[1]) if len(sys.argv) > 1 else 128
piter = int(sys.argv[2]) if len(sys.argv) > 2 else 6
nmodes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
w = workloads.synth_angles(1, 4 * ndet, 4 * ndet, ndet, ndet, 32, nmodes)
psi, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
with pt.CGPtychoSolver(1024, ndet, ndet, 1, 4 * ndet, 4 * ndet) as slv:
    data = slv.fwd(psi, scan, probe[:, 0]).abs().square_().contiguous()
    for k in range(1, nmodes):
        data += slv.fwd(psi, scan, probe[:, k].contiguous()).abs().square_()
    with contextlib.redirect_stdout(io.StringIO()):
        slv.run(data, torch.ones_like(psi), scan.clone(), probe.clone(), piter=piter, recover_prb=True)
    torch.cuda.synchronize()
print("done")
