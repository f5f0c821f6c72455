"""cProfile of the host side of CGPtychoSolver.run (one c2 angle): where does Python wait?"""
import contextlib, cProfile, io, os, pstats, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "libtike-cufft_b200")]
import libtike.cufft as pt
import workloads
ndet = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = workloads.synth_angles(1, 4 * ndet, 4 * ndet, ndet, ndet, 32, 1)
psi, scan, probe = (torch.from_numpy(w[k]).cuda() for k in ("psi", "scan", "probe"))
with pt.CGPtychoSolver(1024, ndet, ndet, 1, 4 * ndet, 4 * ndet) as slv:
    data = slv.fwd(psi, scan, probe[:, 0]).abs().square_().contiguous()
    with contextlib.redirect_stdout(io.StringIO()):
        slv.run(data, torch.ones_like(psi), scan.clone(), probe.clone(), piter=4, recover_prb=True)
        pr = cProfile.Profile()
        pr.enable()
        slv.run(data, torch.ones_like(psi), scan.clone(), probe.clone(), piter=16, recover_prb=True)
        torch.cuda.synchronize()
        pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
